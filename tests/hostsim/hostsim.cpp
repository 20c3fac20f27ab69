// Host simulator (TEST INFRASTRUCTURE ONLY): runs the exact per-thread / per-warp device code of
// parallel-in-time-ode-filters_b200/csrc on the CPU, with the same chunking, tree schedule and memory layout the
// CUDA kernels use, so the math can be checked against the oracle without a GPU.  Never loaded by the product.
#include <cstring>
#include <vector>

#include "pof_pipeline.cuh"

using namespace pof;

template <int d, int q>
static int run(long N, long L, const double* qL, const double* x0, const double* H, const double* c, double* means,
               double* chols, double* fmeans, double* fchols, int calibrate, double* scalars) {
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
  // up-sweep
  for (int l = 0; l + 1 < tl.nlev; ++l)
    for (long i = 0; i < tl.sz[l + 1]; ++i) {
      double* par = &fagg[(tl.off[l + 1] + i) * FE];
      const double* lc = &fagg[(tl.off[l] + 2 * i) * FE];
      if (2 * i + 1 < tl.sz[l]) filter_combine(w, D, lc, lc + FE, par, smem.data(), false);
      else std::memcpy(par, lc, FE * sizeof(double));
    }
  // down-sweep (exclusive, state form)
  std::memcpy(&fin[(tl.off[tl.nlev - 1]) * ST], x0, ST * sizeof(double));
  for (int l = tl.nlev - 1; l >= 1; --l)
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
  for (int l = 0; l + 1 < tl.nlev; ++l)
    for (long i = 0; i < tl.sz[l + 1]; ++i) {
      double* par = &sagg[(tl.off[l + 1] + i) * SE];
      const double* lc = &sagg[(tl.off[l] + 2 * i) * SE];
      if (2 * i + 1 < tl.sz[l]) smooth_combine(w, D, lc + SE, lc, par, smem.data(), false);
      else std::memcpy(par, lc, SE * sizeof(double));
    }
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
  const double cscale = calibrate ? sqrt(ssq) : 1.0;
  for (long ch = 0; ch < CS; ++ch)
    CK::smooth(ch * L, std::min((ch + 1) * L, n), ch == CS - 1, true, qL, &sin_[ch * ST], kern.data(), CS, ch, cscale,
               means, chols, &part2[ch * 2]);
  double obj = 0, bad = 0;
  for (long ch = 0; ch < CS; ++ch) { obj += part2[ch * 2]; bad += part2[ch * 2 + 1]; }
  scalars[0] = nll; scalars[1] = obj; scalars[2] = ssq; scalars[3] = ssqp; scalars[4] = bad;
  return 0;
}

extern "C" {
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
}
