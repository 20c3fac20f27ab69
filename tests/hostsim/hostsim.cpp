// Host simulator (TEST INFRASTRUCTURE ONLY): runs the exact per-thread / per-warp device code of
// parallel-in-time-ode-filters_b200/csrc on the CPU, with the same chunking, tree schedule and memory layout the
// CUDA kernels use, so the math can be checked against the oracle without a GPU.  Never loaded by the product.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "pof_pipeline.cuh"
#include "pof_tile.cuh"

using namespace pof;

// ks_max > 0: the HYBRID tree schedule of the CUDA library (pof_api.cu: flow_hybrid_up / flow_hybrid_down /
// flow_hybrid_suffix, pof_tree_kernels.cuh: KS / KS_APPLY items) -- up-sweep to the first level with <= ks_max nodes,
// Kogge-Stone scan over its nodes (a node's window is final after step nbits(r) of its scan position r and is never
// copied), states / "later" aggregates of that level, down-sweep from there; executed sequentially in item order.
static int hs_nbits(long j) { int b = 0; while (j) { ++b; j >>= 1; } return b; }
template <int d, int q>
static int run(long N, long L, const double* qL, const double* x0, const double* H, const double* c, double* means,
               double* chols, double* fmeans, double* fchols, int calibrate, double* scalars, long ks_max = 0) {
  using CK = Chunk<d, q>;
  constexpr int D = CK::D;
  const long n = N - 1;
  const long CS = (n + L - 1) / L;
  const int FE = 3 * D * D + 2 * D, SE = 2 * D * D + D, ST = D * D + D;
  TreeLevels tl;
  tl.build(CS);
  std::vector<double> fagg(tl.total * FE), fin(tl.total * ST), sagg(tl.total * SE), sin_(tl.total * ST);
  std::vector<double> kern((size_t)L * CK::NE * CS), send(CS * ST), part(CS * 3), part2(CS * 2);
  std::vector<double> smem(coop_ws_doubles(D));
  Warp w;
  // phase 1
  for (long ch = 0; ch < CS; ++ch) CK::fold(ch * L, std::min((ch + 1) * L, n), H, c, qL, &fagg[ch * FE]);
  // base level of the Kogge-Stone stage (hybrid schedule), as WsLayout::build in pof_api.cu
  int B = 0;
  long nB = 0;
  int K = 0;
  if (ks_max > 0) {
    while (B < tl.nlev - 1 && tl.sz[B] > ks_max) ++B;
    nB = tl.sz[B];
    while ((1L << K) < nB) ++K;
    if (tl.nlev - 1 - B < 2) K = 0;
  }
  const int up_top = K > 0 ? B : tl.nlev - 1;
  // up-sweep
  for (int l = 0; l + 1 <= up_top; ++l)
    for (long i = 0; i < tl.sz[l + 1]; ++i) {
      double* par = &fagg[(tl.off[l + 1] + i) * FE];
      const double* lc = &fagg[(tl.off[l] + 2 * i) * FE];
      if (2 * i + 1 < tl.sz[l]) filter_combine(w, D, lc, lc + FE, par, smem.data(), false);
      else std::memcpy(par, lc, FE * sizeof(double));
    }
  int down_from = tl.nlev - 1;
  if (K > 0) {
    // Kogge-Stone PREFIX scan over the nodes of level B, then incoming states of that level from x0
    std::vector<double> ks((size_t)K * nB * FE);
    auto elem = [&](int l, long j) { return l == 0 ? &fagg[(tl.off[B] + j) * FE] : &ks[((size_t)(l - 1) * nB + j) * FE]; };
    for (int st = 1; st <= K; ++st)
      for (long i = 0; i + (1L << (st - 1)) < nB; ++i) {
        const long j = i + (1L << (st - 1));
        filter_combine(w, D, elem(std::min(st - 1, hs_nbits(i)), i), elem(st - 1, j), elem(st, j), smem.data(), false);
      }
    std::memcpy(&fin[tl.off[B] * ST], x0, ST * sizeof(double));
    for (long j = 1; j < nB; ++j)
      filter_combine(w, D, x0, elem(hs_nbits(j - 1), j - 1), &fin[(tl.off[B] + j) * ST], smem.data(), true);
    down_from = B;
  } else {
    std::memcpy(&fin[(tl.off[tl.nlev - 1]) * ST], x0, ST * sizeof(double));
  }
  // down-sweep (exclusive, state form)
  for (int l = down_from; l >= 1; --l)
    for (long i = 0; i < tl.sz[l]; ++i) {
      const double* pin = &fin[(tl.off[l] + i) * ST];
      std::memcpy(&fin[(tl.off[l - 1] + 2 * i) * ST], pin, ST * sizeof(double));
      if (2 * i + 1 < tl.sz[l - 1])
        filter_combine(w, D, pin, &fagg[(tl.off[l - 1] + 2 * i) * FE], &fin[(tl.off[l - 1] + 2 * i + 1) * ST],
                       smem.data(), true);
    }
  if (fmeans) {
    std::memcpy(fmeans, x0, D * sizeof(double));
    std::memcpy(fchols, x0 + D, D * D * sizeof(double));
  }
  // phase 3
  for (long ch = 0; ch < CS; ++ch)
    CK::scan(ch * L, std::min((ch + 1) * L, n), H, c, qL, &fin[ch * ST], kern.data(), CS, ch, &sagg[ch * SE],
             &send[ch * ST], &part[ch * 3], fmeans, fchols);
  double nll = 0, s1 = 0, s2 = 0;
  for (long ch = 0; ch < CS; ++ch) { nll += part[ch * 3]; s1 += part[ch * 3 + 1]; s2 += part[ch * 3 + 2]; }
  const double ssq = s1 / n / d, ssqp = s2 / n / d;
  // smoother tree
  for (int l = 0; l + 1 <= up_top; ++l)
    for (long i = 0; i < tl.sz[l + 1]; ++i) {
      double* par = &sagg[(tl.off[l + 1] + i) * SE];
      const double* lc = &sagg[(tl.off[l] + 2 * i) * SE];
      if (2 * i + 1 < tl.sz[l]) smooth_combine(w, D, lc + SE, lc, par, smem.data(), false);
      else std::memcpy(par, lc, SE * sizeof(double));
    }
  if (K > 0) {
    // element-form SUFFIX scan: Kogge-Stone over level B in scan order r = nB-1-node, "everything later" aggregates,
    // element-form down-sweep, then one state-form combine per chunk with the terminal state
    std::vector<double> ks((size_t)K * nB * SE), sx((size_t)tl.total * SE);
    auto elem = [&](int l, long j) { return l == 0 ? &sagg[(tl.off[B] + j) * SE] : &ks[((size_t)(l - 1) * nB + j) * SE]; };
    for (int st = 1; st <= K; ++st)
      for (long i = 0; i + (1L << (st - 1)) < nB; ++i) {
        const long ja = nB - 1 - i, j = nB - 1 - (i + (1L << (st - 1)));
        smooth_combine(w, D, elem(std::min(st - 1, hs_nbits(i)), ja), elem(st - 1, j), elem(st, j), smem.data(), false);
      }
    for (long i = 0; i < nB; ++i) {
      double* out = &sx[(tl.off[B] + i) * SE];
      const long r = nB - 1 - i;
      if (r == 0) {
        std::fill(out, out + SE, 0.0);
        for (int k = 0; k < D; ++k) out[D + k * D + k] = 1.0;  // identity element: g = 0, E = I, D = 0
      } else {
        std::memcpy(out, elem(hs_nbits(r - 1), i + 1), SE * sizeof(double));
      }
    }
    for (int l = B; l >= 1; --l)
      for (long i = 0; i < tl.sz[l]; ++i) {
        const double* px = &sx[(tl.off[l] + i) * SE];
        if (2 * i + 1 < tl.sz[l - 1]) {
          std::memcpy(&sx[(tl.off[l - 1] + 2 * i + 1) * SE], px, SE * sizeof(double));
          smooth_combine(w, D, px, &sagg[(tl.off[l - 1] + 2 * i + 1) * SE], &sx[(tl.off[l - 1] + 2 * i) * SE],
                         smem.data(), false);
        } else {
          std::memcpy(&sx[(tl.off[l - 1] + 2 * i) * SE], px, SE * sizeof(double));
        }
      }
    for (long ch = 0; ch < CS; ++ch)
      smooth_combine(w, D, &send[(CS - 1) * ST], &sx[ch * SE], &sin_[ch * ST], smem.data(), true);
  } else {
    std::memcpy(&sin_[(tl.off[tl.nlev - 1]) * ST], &send[(CS - 1) * ST], ST * sizeof(double));
    for (int l = tl.nlev - 1; l >= 1; --l)
      for (long i = 0; i < tl.sz[l]; ++i) {
        const double* pin = &sin_[(tl.off[l] + i) * ST];
        if (2 * i + 1 < tl.sz[l - 1]) {
          std::memcpy(&sin_[(tl.off[l - 1] + 2 * i + 1) * ST], pin, ST * sizeof(double));
          smooth_combine(w, D, pin, &sagg[(tl.off[l - 1] + 2 * i + 1) * SE], &sin_[(tl.off[l - 1] + 2 * i) * ST],
                         smem.data(), true);
        } else {
          std::memcpy(&sin_[(tl.off[l - 1] + 2 * i) * ST], pin, ST * sizeof(double));
        }
      }
  }
  const double cscale = calibrate ? sqrt(ssq) : 1.0;
  for (long ch = 0; ch < CS; ++ch)
    CK::smooth(ch * L, std::min((ch + 1) * L, n), ch == CS - 1, true, qL, &sin_[ch * ST], kern.data(), CS, ch, cscale,
               means, chols, &part2[ch * 2]);
  double obj = 0, bad = 0;
  for (long ch = 0; ch < CS; ++ch) { obj += part2[ch * 2]; bad += part2[ch * 2 + 1]; }
  scalars[0] = nll; scalars[1] = obj; scalars[2] = ssq; scalars[3] = ssqp; scalars[4] = bad;
  return 0;
}

// ---- time-sharded form: the same three stages the CUDA library exposes (pof_shard_stage_{a,b,c}_f64), on the host
struct HsWs {
  int d, q, D, FE, SE, ST;
  long n, L, CS;
  TreeLevels tl;
  std::vector<double> fagg, fin, sagg, sin_, kern, send, part, part2, smem, qL;
};
template <int d, int q>
static void hs_a(HsWs& w, const double* H, const double* c, double* carry_f) {
  using CK = Chunk<d, q>;
  Warp wp;
  for (long ch = 0; ch < w.CS; ++ch)
    CK::fold(ch * w.L, std::min((ch + 1) * w.L, w.n), H, c, w.qL.data(), &w.fagg[ch * w.FE]);
  for (int l = 0; l + 1 < w.tl.nlev; ++l)
    for (long i = 0; i < w.tl.sz[l + 1]; ++i) {
      double* par = &w.fagg[(w.tl.off[l + 1] + i) * w.FE];
      const double* lc = &w.fagg[(w.tl.off[l] + 2 * i) * w.FE];
      if (2 * i + 1 < w.tl.sz[l]) filter_combine(wp, w.D, lc, lc + w.FE, par, w.smem.data(), false);
      else std::memcpy(par, lc, w.FE * sizeof(double));
    }
  std::memcpy(carry_f, &w.fagg[w.tl.off[w.tl.nlev - 1] * w.FE], w.FE * sizeof(double));
}
template <int d, int q>
static void hs_b(HsWs& w, const double* H, const double* c, const double* state_in, double* fmeans, double* fchols,
                 double* carry_s, double* state_end, double* partials) {
  using CK = Chunk<d, q>;
  Warp wp;
  const TreeLevels& tl = w.tl;
  std::memcpy(&w.fin[tl.off[tl.nlev - 1] * w.ST], state_in, w.ST * sizeof(double));
  for (int l = tl.nlev - 1; l >= 1; --l)
    for (long i = 0; i < tl.sz[l]; ++i) {
      const double* pin = &w.fin[(tl.off[l] + i) * w.ST];
      std::memcpy(&w.fin[(tl.off[l - 1] + 2 * i) * w.ST], pin, w.ST * sizeof(double));
      if (2 * i + 1 < tl.sz[l - 1])
        filter_combine(wp, w.D, pin, &w.fagg[(tl.off[l - 1] + 2 * i) * w.FE],
                       &w.fin[(tl.off[l - 1] + 2 * i + 1) * w.ST], w.smem.data(), true);
    }
  for (long ch = 0; ch < w.CS; ++ch)
    CK::scan(ch * w.L, std::min((ch + 1) * w.L, w.n), H, c, w.qL.data(), &w.fin[ch * w.ST], w.kern.data(), w.CS, ch,
             &w.sagg[ch * w.SE], &w.send[ch * w.ST], &w.part[ch * 3], fmeans, fchols);
  for (int l = 0; l + 1 < tl.nlev; ++l)
    for (long i = 0; i < tl.sz[l + 1]; ++i) {
      double* par = &w.sagg[(tl.off[l + 1] + i) * w.SE];
      const double* lc = &w.sagg[(tl.off[l] + 2 * i) * w.SE];
      if (2 * i + 1 < tl.sz[l]) smooth_combine(wp, w.D, lc + w.SE, lc, par, w.smem.data(), false);
      else std::memcpy(par, lc, w.SE * sizeof(double));
    }
  std::memcpy(carry_s, &w.sagg[tl.off[tl.nlev - 1] * w.SE], w.SE * sizeof(double));
  std::memcpy(state_end, &w.send[(w.CS - 1) * w.ST], w.ST * sizeof(double));
  partials[0] = partials[1] = partials[2] = 0.0;
  for (long ch = 0; ch < w.CS; ++ch)
    for (int j = 0; j < 3; ++j) partials[j] += w.part[ch * 3 + j];
}
template <int d, int q>
static void hs_c(HsWs& w, const double* seed, int has_row0, double cscale, double* means, double* chols,
                 double* partials2) {
  using CK = Chunk<d, q>;
  Warp wp;
  const TreeLevels& tl = w.tl;
  std::memcpy(&w.sin_[tl.off[tl.nlev - 1] * w.ST], seed, w.ST * sizeof(double));
  for (int l = tl.nlev - 1; l >= 1; --l)
    for (long i = 0; i < tl.sz[l]; ++i) {
      const double* pin = &w.sin_[(tl.off[l] + i) * w.ST];
      if (2 * i + 1 < tl.sz[l - 1]) {
        std::memcpy(&w.sin_[(tl.off[l - 1] + 2 * i + 1) * w.ST], pin, w.ST * sizeof(double));
        smooth_combine(wp, w.D, pin, &w.sagg[(tl.off[l - 1] + 2 * i + 1) * w.SE],
                       &w.sin_[(tl.off[l - 1] + 2 * i) * w.ST], w.smem.data(), true);
      } else {
        std::memcpy(&w.sin_[(tl.off[l - 1] + 2 * i) * w.ST], pin, w.ST * sizeof(double));
      }
    }
  const long shift = has_row0 ? 0 : 1;
  double* mb = means - shift * w.D;
  double* cb = chols ? chols - shift * (long)w.D * w.D : nullptr;
  for (long ch = 0; ch < w.CS; ++ch)
    CK::smooth(ch * w.L, std::min((ch + 1) * w.L, w.n), ch == w.CS - 1, has_row0 != 0, w.qL.data(),
               &w.sin_[ch * w.ST], w.kern.data(), w.CS, ch, cscale, mb, cb, &w.part2[ch * 2]);
  partials2[0] = partials2[1] = 0.0;
  for (long ch = 0; ch < w.CS; ++ch)
    for (int j = 0; j < 2; ++j) partials2[j] += w.part2[ch * 2 + j];
}

#define HS_DISPATCH(CALL)                                                                                      \
  do {                                                                                                         \
    const int d = w->d, q = w->q;                                                                              \
    if (d == 1 && q == 1) { CALL(1, 1); } else if (d == 1 && q == 2) { CALL(1, 2); }                           \
    else if (d == 1 && q == 3) { CALL(1, 3); } else if (d == 2 && q == 2) { CALL(2, 2); }                      \
    else if (d == 2 && q == 3) { CALL(2, 3); } else if (d == 3 && q == 3) { CALL(3, 3); }                      \
    else return -1;                                                                                            \
  } while (0)

extern "C" {
void* hs_ws_create(int d, int q, long n_loc, long L, const double* qL) {
  HsWs* w = new HsWs;
  w->d = d; w->q = q; w->D = d * (q + 1);
  const int D = w->D;
  w->FE = 3 * D * D + 2 * D; w->SE = 2 * D * D + D; w->ST = D * D + D;
  w->n = n_loc; w->L = L; w->CS = (n_loc + L - 1) / L;
  w->tl.build(w->CS);
  w->fagg.resize(w->tl.total * w->FE); w->fin.resize(w->tl.total * w->ST);
  w->sagg.resize(w->tl.total * w->SE); w->sin_.resize(w->tl.total * w->ST);
  w->kern.resize((size_t)L * (D + 2 * D * D) * w->CS); w->send.resize(w->CS * w->ST);
  w->part.resize(w->CS * 3); w->part2.resize(w->CS * 2);
  w->smem.resize(coop_ws_doubles(D));
  w->qL.assign(qL, qL + (q + 1) * (q + 1));
  return w;
}
void hs_ws_free(void* p) { delete (HsWs*)p; }
int hs_stage_a(void* p, const double* H, const double* c, double* carry_f) {
  HsWs* w = (HsWs*)p;
#define CALL(dd, qq) hs_a<dd, qq>(*w, H, c, carry_f)
  HS_DISPATCH(CALL);
#undef CALL
  return 0;
}
int hs_stage_b(void* p, const double* H, const double* c, const double* state_in, double* fmeans, double* fchols,
               double* carry_s, double* state_end, double* partials) {
  HsWs* w = (HsWs*)p;
#define CALL(dd, qq) hs_b<dd, qq>(*w, H, c, state_in, fmeans, fchols, carry_s, state_end, partials)
  HS_DISPATCH(CALL);
#undef CALL
  return 0;
}
int hs_stage_c(void* p, const double* seed, int has_row0, double cscale, double* means, double* chols,
               double* partials2) {
  HsWs* w = (HsWs*)p;
#define CALL(dd, qq) hs_c<dd, qq>(*w, seed, has_row0, cscale, means, chols, partials2)
  HS_DISPATCH(CALL);
#undef CALL
  return 0;
}
int hs_filter_chain(int D, int count, const double* state_in, const double* elems, double* state_out) {
  std::vector<double> smem(coop_ws_doubles(D)), cur(state_in, state_in + D + D * D), nxt(D + D * D);
  Warp w;
  const int FE = 3 * D * D + 2 * D;
  for (int i = 0; i < count; ++i) {
    filter_combine(w, D, cur.data(), elems + (long)i * FE, nxt.data(), smem.data(), true);
    cur.swap(nxt);
  }
  std::memcpy(state_out, cur.data(), (D + D * D) * sizeof(double));
  return 0;
}
int hs_smooth_chain(int D, int count, const double* state_in, const double* elems, double* state_out) {
  std::vector<double> smem(coop_ws_doubles(D)), cur(state_in, state_in + D + D * D), nxt(D + D * D);
  Warp w;
  const int SE = 2 * D * D + D;
  for (int i = 0; i < count; ++i) {
    smooth_combine(w, D, cur.data(), elems + (long)(count - 1 - i) * SE, nxt.data(), smem.data(), true);
    cur.swap(nxt);
  }
  std::memcpy(state_out, cur.data(), (D + D * D) * sizeof(double));
  return 0;
}
int hs_seq_eks(int d, int q, long N, const double* qL, double s0, double s1, int ivp_id, const double* params8,
               const double* x0, double* means, double* chols, double* part) {
  IvpParams P;
  for (int i = 0; i < 8; ++i) P.p[i] = params8[i];
  const int D = d * (q + 1);
  std::vector<double> kern((size_t)(N - 1) * (D + 2 * D * D));
#define CASE(dd, qq)                                                                                          \
  if (d == dd && q == qq) {                                                                                   \
    Chunk<dd, qq>::seq_eks(N - 1, s0, s1, qL, ivp_id, P, x0, kern.data(), means, chols, part);                \
    return 0;                                                                                                 \
  }
  CASE(1, 1) CASE(1, 3) CASE(2, 3) CASE(3, 2)
#undef CASE
  return -1;
}
int hs_filter_combine(int D, const double* e1, const double* e2, double* out, int state_mode) {
  std::vector<double> smem(coop_ws_doubles(D));
  Warp w;
  filter_combine(w, D, e1, e2, out, smem.data(), state_mode != 0);
  return 0;
}
int hs_smooth_combine(int D, const double* e1, const double* e2, double* out, int state_mode) {
  std::vector<double> smem(coop_ws_doubles(D));
  Warp w;
  smooth_combine(w, D, e1, e2, out, smem.data(), state_mode != 0);
  return 0;
}
int hs_linear_filtsmooth(int d, int q, long N, long L, const double* qL, const double* x0, const double* H,
                         const double* c, double* means, double* chols, double* fmeans, double* fchols,
                         int calibrate, double* scalars) {
#define CASE(dd, qq) \
  if (d == dd && q == qq) return run<dd, qq>(N, L, qL, x0, H, c, means, chols, fmeans, fchols, calibrate, scalars);
  CASE(1, 1) CASE(1, 2) CASE(1, 3) CASE(1, 4) CASE(2, 1) CASE(2, 2) CASE(2, 3) CASE(3, 3) CASE(4, 2) CASE(4, 3)
#undef CASE
  return -1;
}
// the same pass with the hybrid tree schedule (ks_max = WsLayout::KS_MAX in the library: 384)
int hs_linear_filtsmooth_hybrid(int d, int q, long N, long L, long ks_max, const double* qL, const double* x0,
                                const double* H, const double* c, double* means, double* chols, double* fmeans,
                                double* fchols, int calibrate, double* scalars) {
#define CASE(dd, qq)          \
  if (d == dd && q == qq)     \
    return run<dd, qq>(N, L, qL, x0, H, c, means, chols, fmeans, fchols, calibrate, scalars, ks_max);
  CASE(1, 1) CASE(1, 2) CASE(1, 3) CASE(1, 4) CASE(2, 1) CASE(2, 2) CASE(2, 3) CASE(3, 3) CASE(4, 2) CASE(4, 3)
#undef CASE
  return -1;
}
}

// ---- large-state ("tile") family: CTA-cooperative device code of pof_tile.cuh, executed by a one-thread team whose
// iteration order inside every Team::each is selectable (0 forward, 1 reverse, 2 permuted): identical results under
// all orders <=> the iterations of every each() are independent <=> the CUDA kernels are race-free (pof_tile.cuh).
static int run_tile(int d, int q, long N, long L, const double* qL, const double* x0, const double* H, const double* c,
                    const double* Jc, double s0, double s1, const double* R, const double* Fd, const double* Qd,
                    double* means, double* chols, double* fmeans, double* fchols, int calibrate, double* scalars,
                    int order, int reg_sweeps) {
  Team::order() = order;
  Team t;
  const int D = d * (q + 1);
  const long n = N - 1;
  const long CS = (n + L - 1) / L;
  const int FE = 3 * D * D + 2 * D, SE = 2 * D * D + D, ST = D * D + D, NE = D + 2 * D * D;
  TreeLevels tl;
  tl.build(CS);
  std::vector<double> fagg(tl.total * FE), faggm(CS * FE), fin(tl.total * ST), sagg(tl.total * SE), sin_(tl.total * ST);
  std::vector<double> kern((size_t)n * NE), send(CS * ST), part(CS * 3), part2(CS * 2);
  std::vector<double> smem(std::max({tile_fold_smem_doubles(D, d), tile_scan_smem_doubles(D, d),
                                     tile_smooth_smem_doubles(D, d), tile_tree_smem_doubles(D)}));
  // poison the shared memory so that reads of never-written entries show up as NaN
  auto poison = [&]() { std::fill(smem.begin(), smem.end(), std::nan("")); };
  const TileLin lin = {H, c, Jc, R, s0, s1, Fd, Qd, reg_sweeps};
  for (long ch = 0; ch < CS; ++ch) {
    poison();
    tile_fold(t, d, q, qL, lin, ch * L, std::min((ch + 1) * L, n), &fagg[ch * FE], &faggm[ch * FE], smem.data());
  }
  for (int l = 0; l + 1 < tl.nlev; ++l)
    for (long i = 0; i < tl.sz[l + 1]; ++i) {
      double* par = &fagg[(tl.off[l + 1] + i) * FE];
      const double* lc = &fagg[(tl.off[l] + 2 * i) * FE];
      poison();
      if (2 * i + 1 < tl.sz[l]) tile_filter_combine(t, D, lc, lc + FE, par, smem.data(), false);
      else std::memcpy(par, lc, FE * sizeof(double));
    }
  std::memcpy(&fin[(tl.off[tl.nlev - 1]) * ST], x0, ST * sizeof(double));
  for (int l = tl.nlev - 1; l >= 1; --l)
    for (long i = 0; i < tl.sz[l]; ++i) {
      const double* pin = &fin[(tl.off[l] + i) * ST];
      std::memcpy(&fin[(tl.off[l - 1] + 2 * i) * ST], pin, ST * sizeof(double));
      poison();
      if (2 * i + 1 < tl.sz[l - 1])
        tile_filter_combine(t, D, pin, &fagg[(tl.off[l - 1] + 2 * i) * FE], &fin[(tl.off[l - 1] + 2 * i + 1) * ST],
                            smem.data(), true);
    }
  if (fmeans) {
    std::memcpy(fmeans, x0, D * sizeof(double));
    std::memcpy(fchols, x0 + D, D * D * sizeof(double));
  }
  for (long ch = 0; ch < CS; ++ch) {
    poison();
    tile_chunk_kernel(t, D, &fin[ch * ST], &faggm[ch * FE], &sagg[ch * SE], smem.data());
    poison();
    tile_scan(t, d, q, qL, lin, ch * L, std::min((ch + 1) * L, n), &fin[ch * ST], kern.data(), &send[ch * ST],
              &part[ch * 3], fmeans, fchols, smem.data());
  }
  double nll = 0, a1 = 0, a2 = 0;
  for (long ch = 0; ch < CS; ++ch) { nll += part[ch * 3]; a1 += part[ch * 3 + 1]; a2 += part[ch * 3 + 2]; }
  const double ssq = a1 / n / d, ssqp = a2 / n / d;
  for (int l = 0; l + 1 < tl.nlev; ++l)
    for (long i = 0; i < tl.sz[l + 1]; ++i) {
      double* par = &sagg[(tl.off[l + 1] + i) * SE];
      const double* lc = &sagg[(tl.off[l] + 2 * i) * SE];
      poison();
      if (2 * i + 1 < tl.sz[l]) tile_smooth_combine(t, D, lc + SE, lc, par, smem.data(), false);
      else std::memcpy(par, lc, SE * sizeof(double));
    }
  std::memcpy(&sin_[(tl.off[tl.nlev - 1]) * ST], &send[(CS - 1) * ST], ST * sizeof(double));
  for (int l = tl.nlev - 1; l >= 1; --l)
    for (long i = 0; i < tl.sz[l]; ++i) {
      const double* pin = &sin_[(tl.off[l] + i) * ST];
      if (2 * i + 1 < tl.sz[l - 1]) {
        std::memcpy(&sin_[(tl.off[l - 1] + 2 * i + 1) * ST], pin, ST * sizeof(double));
        poison();
        tile_smooth_combine(t, D, pin, &sagg[(tl.off[l - 1] + 2 * i + 1) * SE], &sin_[(tl.off[l - 1] + 2 * i) * ST],
                            smem.data(), true);
      } else {
        std::memcpy(&sin_[(tl.off[l - 1] + 2 * i) * ST], pin, ST * sizeof(double));
      }
    }
  const double cscale = calibrate ? sqrt(ssq) : 1.0;
  for (long ch = 0; ch < CS; ++ch) {
    poison();
    tile_smooth(t, d, q, qL, lin, ch * L, std::min((ch + 1) * L, n), ch == CS - 1, true, &sin_[ch * ST], kern.data(), cscale,
                means, chols, &part2[ch * 2], smem.data());
  }
  double obj = 0, bad = 0;
  for (long ch = 0; ch < CS; ++ch) { obj += part2[ch * 2]; bad += part2[ch * 2 + 1]; }
  scalars[0] = nll; scalars[1] = obj; scalars[2] = ssq; scalars[3] = ssqp; scalars[4] = bad;
  Team::order() = 0;
  return 0;
}

extern "C" {
int hs_tile_linear_filtsmooth(int d, int q, long N, long L, const double* qL, const double* x0, const double* H,
                              const double* c, const double* Jc, double s0, double s1, const double* R,
                              const double* Fd, const double* Qd, double* means, double* chols, double* fmeans,
                              double* fchols, int calibrate, double* scalars, int order, int reg_sweeps) {
  return run_tile(d, q, N, L, qL, x0, H, c, Jc, s0, s1, R, Fd, Qd, means, chols, fmeans, fchols, calibrate, scalars,
                  order, reg_sweeps);
}
int hs_tile_filter_combine(int D, const double* e1, const double* e2, double* out, int state_mode, int order) {
  std::vector<double> smem(tile_tree_smem_doubles(D), std::nan(""));
  Team::order() = order;
  Team t;
  tile_filter_combine(t, D, e1, e2, out, smem.data(), state_mode != 0);
  Team::order() = 0;
  return 0;
}
int hs_tile_smooth_combine(int D, const double* e1, const double* e2, double* out, int state_mode, int order) {
  std::vector<double> smem(tile_tree_smem_doubles(D), std::nan(""));
  Team::order() = order;
  Team t;
  tile_smooth_combine(t, D, e1, e2, out, smem.data(), state_mode != 0);
  Team::order() = 0;
  return 0;
}
int hs_tile_seq_eks(int d, int q, long N, const double* qL, double s0, double s1, int ivp_id, const double* params8,
                    const double* x0, double* means, double* chols, double* sums, int order, int reg_sweeps) {
  Team::order() = order;
  Team t;
  const int D = d * (q + 1);
  TileEks eks;
  eks.ivp_id = ivp_id;
  for (int i = 0; i < 8; ++i) eks.P.p[i] = params8[i];
  const TileLin lin = {nullptr, nullptr, nullptr, nullptr, s0, s1, nullptr, nullptr, reg_sweeps};
  std::vector<double> kern((size_t)(N - 1) * (D + 2 * D * D)), send(D + D * D);
  std::vector<double> smem(std::max(tile_scan_smem_doubles(D, d), tile_smooth_smem_doubles(D, d)), std::nan(""));
  tile_seq_eks(t, d, q, qL, lin, eks, N - 1, x0, kern.data(), send.data(), means, chols, sums, smem.data());
  Team::order() = 0;
  return 0;
}
// Lorenz-96 linearisation (the body of k_linearize_l96): dense (H, c) and compact [J_f | c]
int hs_linearize_l96(double forcing, long n, int d, int q, double s0, double s1, const double* means_t1, double* H,
                     double* c, double* Jc) {
  for (long k = 0; k < n; ++k)
    for (int a = 0; a < d; ++a) {
      l96_linearize_row(forcing, k, a, d, q, s0, s1, 1, means_t1, H, c, nullptr);
      l96_linearize_row(forcing, k, a, d, q, s0, 0.0, 0, means_t1, nullptr, nullptr, Jc);
    }
  return 0;
}
// one Householder sweep (register or shared-memory version) on a caller-supplied array: unit test hook
int hs_tile_tria(double* M, int R, int C, int ld, int npiv, int c0, int use_reg, int order) {
  std::vector<double> diag(4 * (R + C) + 16), pb(2 * TILE_PB_COLS_);
  Team::order() = order;
  Team t;
  tile_tria(t, M, R, C, ld, npiv, c0, diag.data(), use_reg ? pb.data() : nullptr);
  Team::order() = 0;
  return 0;
}
int hs_tile_smem_bytes(int D, int d, int which) {
  const int v[4] = {tile_fold_smem_doubles(D, d), tile_scan_smem_doubles(D, d), tile_smooth_smem_doubles(D, d),
                    tile_tree_smem_doubles(D)};
  return v[which & 3] * (int)sizeof(double);
}
}

// ---- time-sharded form on the tile family: the three stages of pof_shard_stage_{a,b,c}_f64 when the leaves are the
// CTA-per-chunk kernels and the tree is the CTA-per-node one (runtime d, q)
struct HsTileWs {
  int d, q, D, FE, SE, ST, NE;
  long n, L, CS;
  TreeLevels tl;
  std::vector<double> fagg, faggm, fin, sagg, sin_, kern, send, part, part2, smem, qL;
};
extern "C" {
void* hs_tile_ws_create(int d, int q, long n_loc, long L, const double* qL) {
  HsTileWs* w = new HsTileWs;
  w->d = d; w->q = q; w->D = d * (q + 1);
  const int D = w->D;
  w->FE = 3 * D * D + 2 * D; w->SE = 2 * D * D + D; w->ST = D * D + D; w->NE = D + 2 * D * D;
  w->n = n_loc; w->L = L; w->CS = (n_loc + L - 1) / L;
  w->tl.build(w->CS);
  w->fagg.resize(w->tl.total * w->FE); w->faggm.resize(w->CS * w->FE); w->fin.resize(w->tl.total * w->ST);
  w->sagg.resize(w->tl.total * w->SE); w->sin_.resize(w->tl.total * w->ST);
  w->kern.resize((size_t)n_loc * w->NE); w->send.resize(w->CS * w->ST);
  w->part.resize(w->CS * 3); w->part2.resize(w->CS * 2);
  w->smem.resize(std::max({tile_fold_smem_doubles(D, d), tile_scan_smem_doubles(D, d), tile_smooth_smem_doubles(D, d),
                           tile_tree_smem_doubles(D)}));
  w->qL.assign(qL, qL + (q + 1) * (q + 1));
  return w;
}
void hs_tile_ws_free(void* p) { delete (HsTileWs*)p; }
int hs_tile_stage_a(void* p, const double* H, const double* c, double* carry_f) {
  HsTileWs& w = *(HsTileWs*)p;
  Team t;
  const TileLin lin = {H, c, nullptr, nullptr, 0.0, 0.0, nullptr, nullptr, 1};
  for (long ch = 0; ch < w.CS; ++ch)
    tile_fold(t, w.d, w.q, w.qL.data(), lin, ch * w.L, std::min((ch + 1) * w.L, w.n), &w.fagg[ch * w.FE],
              &w.faggm[ch * w.FE], w.smem.data());
  for (int l = 0; l + 1 < w.tl.nlev; ++l)
    for (long i = 0; i < w.tl.sz[l + 1]; ++i) {
      double* par = &w.fagg[(w.tl.off[l + 1] + i) * w.FE];
      const double* lc = &w.fagg[(w.tl.off[l] + 2 * i) * w.FE];
      if (2 * i + 1 < w.tl.sz[l]) tile_filter_combine(t, w.D, lc, lc + w.FE, par, w.smem.data(), false);
      else std::memcpy(par, lc, w.FE * sizeof(double));
    }
  std::memcpy(carry_f, &w.fagg[w.tl.off[w.tl.nlev - 1] * w.FE], w.FE * sizeof(double));
  return 0;
}
int hs_tile_stage_b(void* p, const double* H, const double* c, const double* state_in, double* fmeans, double* fchols,
                    double* carry_s, double* state_end, double* partials) {
  HsTileWs& w = *(HsTileWs*)p;
  Team t;
  const TreeLevels& tl = w.tl;
  const TileLin lin = {H, c, nullptr, nullptr, 0.0, 0.0, nullptr, nullptr, 1};
  std::memcpy(&w.fin[tl.off[tl.nlev - 1] * w.ST], state_in, w.ST * sizeof(double));
  for (int l = tl.nlev - 1; l >= 1; --l)
    for (long i = 0; i < tl.sz[l]; ++i) {
      const double* pin = &w.fin[(tl.off[l] + i) * w.ST];
      std::memcpy(&w.fin[(tl.off[l - 1] + 2 * i) * w.ST], pin, w.ST * sizeof(double));
      if (2 * i + 1 < tl.sz[l - 1])
        tile_filter_combine(t, w.D, pin, &w.fagg[(tl.off[l - 1] + 2 * i) * w.FE],
                            &w.fin[(tl.off[l - 1] + 2 * i + 1) * w.ST], w.smem.data(), true);
    }
  for (long ch = 0; ch < w.CS; ++ch) {
    tile_chunk_kernel(t, w.D, &w.fin[ch * w.ST], &w.faggm[ch * w.FE], &w.sagg[ch * w.SE], w.smem.data());
    tile_scan(t, w.d, w.q, w.qL.data(), lin, ch * w.L, std::min((ch + 1) * w.L, w.n), &w.fin[ch * w.ST], w.kern.data(),
              &w.send[ch * w.ST], &w.part[ch * 3], fmeans, fchols, w.smem.data());
  }
  for (int l = 0; l + 1 < tl.nlev; ++l)
    for (long i = 0; i < tl.sz[l + 1]; ++i) {
      double* par = &w.sagg[(tl.off[l + 1] + i) * w.SE];
      const double* lc = &w.sagg[(tl.off[l] + 2 * i) * w.SE];
      if (2 * i + 1 < tl.sz[l]) tile_smooth_combine(t, w.D, lc + w.SE, lc, par, w.smem.data(), false);
      else std::memcpy(par, lc, w.SE * sizeof(double));
    }
  std::memcpy(carry_s, &w.sagg[tl.off[tl.nlev - 1] * w.SE], w.SE * sizeof(double));
  std::memcpy(state_end, &w.send[(w.CS - 1) * w.ST], w.ST * sizeof(double));
  partials[0] = partials[1] = partials[2] = 0.0;
  for (long ch = 0; ch < w.CS; ++ch)
    for (int j = 0; j < 3; ++j) partials[j] += w.part[ch * 3 + j];
  return 0;
}
int hs_tile_stage_c(void* p, const double* seed, int has_row0, double cscale, double* means, double* chols,
                    double* partials2) {
  HsTileWs& w = *(HsTileWs*)p;
  Team t;
  const TreeLevels& tl = w.tl;
  const TileLin lin = {nullptr, nullptr, nullptr, nullptr, 0.0, 0.0, nullptr, nullptr, 1};
  std::memcpy(&w.sin_[tl.off[tl.nlev - 1] * w.ST], seed, w.ST * sizeof(double));
  for (int l = tl.nlev - 1; l >= 1; --l)
    for (long i = 0; i < tl.sz[l]; ++i) {
      const double* pin = &w.sin_[(tl.off[l] + i) * w.ST];
      if (2 * i + 1 < tl.sz[l - 1]) {
        std::memcpy(&w.sin_[(tl.off[l - 1] + 2 * i + 1) * w.ST], pin, w.ST * sizeof(double));
        tile_smooth_combine(t, w.D, pin, &w.sagg[(tl.off[l - 1] + 2 * i + 1) * w.SE],
                            &w.sin_[(tl.off[l - 1] + 2 * i) * w.ST], w.smem.data(), true);
      } else {
        std::memcpy(&w.sin_[(tl.off[l - 1] + 2 * i) * w.ST], pin, w.ST * sizeof(double));
      }
    }
  const long shift = has_row0 ? 0 : 1;
  double* mb = means - shift * w.D;
  double* cb = chols ? chols - shift * (long)w.D * w.D : nullptr;
  for (long ch = 0; ch < w.CS; ++ch)
    tile_smooth(t, w.d, w.q, w.qL.data(), lin, ch * w.L, std::min((ch + 1) * w.L, w.n), ch == w.CS - 1, has_row0 != 0,
                &w.sin_[ch * w.ST], w.kern.data(), cscale, mb, cb, &w.part2[ch * 2], w.smem.data());
  partials2[0] = partials2[1] = 0.0;
  for (long ch = 0; ch < w.CS; ++ch)
    for (int j = 0; j < 2; ++j) partials2[j] += w.part2[ch * 2 + j];
  return 0;
}
// the bodies of k_tile_fchain / k_tile_schain
int hs_tile_filter_chain(int D, int count, const double* state_in, const double* elems, double* state_out) {
  std::vector<double> smem(tile_tree_smem_doubles(D)), cur(state_in, state_in + D + D * D), nxt(D + D * D);
  Team t;
  const int FE = 3 * D * D + 2 * D;
  for (int i = 0; i < count; ++i) {
    tile_filter_combine(t, D, cur.data(), elems + (long)i * FE, nxt.data(), smem.data(), true);
    cur.swap(nxt);
  }
  std::memcpy(state_out, cur.data(), (D + D * D) * sizeof(double));
  return 0;
}
int hs_tile_smooth_chain(int D, int count, const double* state_in, const double* elems, double* state_out) {
  std::vector<double> smem(tile_tree_smem_doubles(D)), cur(state_in, state_in + D + D * D), nxt(D + D * D);
  Team t;
  const int SE = 2 * D * D + D;
  for (int i = 0; i < count; ++i) {
    tile_smooth_combine(t, D, cur.data(), elems + (long)(count - 1 - i) * SE, nxt.data(), smem.data(), true);
    cur.swap(nxt);
  }
  std::memcpy(state_out, cur.data(), (D + D * D) * sizeof(double));
  return 0;
}
}
