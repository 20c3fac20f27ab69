// C ABI (include/pof_b200.h) of the B200-native parallel-in-time IEKS pass, plus the kernels that are not templated on
// (d, q): deterministic scalar reductions, the fused vector-field/Jacobian linearisation for the built-in IVPs, the
// prior initial trajectory, the final calibration + projection, and the orchestration of one pass:
//
//   fold (leaf)  ->  filter tree (ONE dataflow launch: up-sweep + down-sweep)  ->  scan (leaf)  ||  smoother up-sweep
//   (side stream of the caller's context)  ->  smoother down-sweep (one dataflow launch)  ->  smooth (leaf)
//
// Kernel families: `lane2` (pof_lane2.cuh: G lanes per chunk, two rows per lane; d <= 4, D <= 16) with the
// register-resident tree operators (pof_treelane.cuh), and `tile` (pof_tile.cuh: one CTA per chunk / tree node; any
// (d, q) whose tiles fit shared memory, observation noise, general per-step models).  There is no other path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "../../include/pof_b200.h"
#include "pof_ctx.h"
#include "pof_ivp.cuh"
#include "pof_launch.cuh"
#include "pof_tree_levels.cuh"
#ifndef POF_F32
#include "pof_coop.cuh"
#include "pof_pipeline.cuh"
#endif

namespace POF_NS {
using namespace pof;  // IvpParams, ivp_eval, TreeLevels (fp64-only helpers shared by both builds)

static const TreeLaunch* tree_launch(int D) {
  const TreeLaunch* t = tree_launch_a(D);
  if (!t) t = tree_launch_b(D);
  if (!t) t = tree_launch_c(D);
  return t;
}
// two rows per lane (pof_lane2.cuh): instantiated for d <= 4, D <= 16
static const LeafLaunch* lane2_launch(int d, int q) {
  switch (d) {
    case 1: return lane2_launch_d1(q);
    case 2: return lane2_launch_d2(q);
    case 3: return lane2_launch_d3(q);
    case 4: return lane2_launch_d4(q);
    default: return nullptr;
  }
}
#ifndef POF_F32
// one-thread sequential EKS (pof_seq_kernels.cuh), d <= 4
static const LeafLaunch* seq_launch(int d, int q) {
  switch (d) {
    case 1: return seq_launch_d1(q);
    case 2: return seq_launch_d2(q);
    case 3: return seq_launch_d3(q);
    case 4: return seq_launch_d4(q);
    default: return nullptr;
  }
}
#endif  // !POF_F32
// the kernel family that serves (d, q): lane2 where instantiated, else (or when POF_F_FAMILY_TILE asks for it) tile
static const LeafLaunch* leaf_launch(int d, int q, unsigned flags) {
  if (!(flags & POF_F_FAMILY_TILE)) {
    if (const LeafLaunch* l2 = lane2_launch(d, q)) return l2;
  }
  return tile_supported(d, q) ? tile_leaf_launch() : nullptr;
}
// tree operators that go with a leaf family: register-resident (with the lane2 leaves) or CTA-per-node tiles
static const TreeLaunch* tree_for(const LeafLaunch* ll, int D, unsigned flags) {
  if (ll->is_tile || (flags & POF_F_FAMILY_TILE)) return nullptr;
  return tree_launch(D);
}

#ifndef POF_F32
// ------------------------------------------------------------------------------------------------ rank-carry chains
// sequential chains over a handful of rank carries (one warp)
__global__ void __launch_bounds__(32)
    k_filter_chain(int D, int count, const real* __restrict__ state_in, const real* __restrict__ elems,
                   real* __restrict__ state_out, real* __restrict__ scratch) {
  extern __shared__ real sm[];
  Warp w;
  const int FE = filter_elem_size(D), ST = state_size(D);
  // ping-pong between state_out and scratch so that the last write lands in state_out
  const real* cur = state_in;
  for (int i = 0; i < count; ++i) {
    real* dst = ((count - 1 - i) % 2 == 0) ? state_out : scratch;
    filter_combine(w, D, cur, elems + (long)i * FE, dst, sm, true);
    w.sync();
    cur = dst;
  }
  if (count == 0) coop_copy(w, state_out, state_in, ST);
}
__global__ void __launch_bounds__(32)
    k_smooth_chain(int D, int count, const real* __restrict__ state_in, const real* __restrict__ elems,
                   real* __restrict__ state_out, real* __restrict__ scratch) {
  extern __shared__ real sm[];
  Warp w;
  const int SE = smooth_elem_size(D), ST = state_size(D);
  const real* cur = state_in;
  for (int i = 0; i < count; ++i) {
    real* dst = ((count - 1 - i) % 2 == 0) ? state_out : scratch;
    smooth_combine(w, D, cur, elems + (long)(count - 1 - i) * SE, dst, sm, true);
    w.sync();
    cur = dst;
  }
  if (count == 0) coop_copy(w, state_out, state_in, ST);
}


#endif  // !POF_F32
// ------------------------------------------------------------------------------------------------ reductions
// out[j] = sum_i part[i*NC + j], fixed summation order (deterministic), single CTA of 1024 threads: every thread
// accumulates its strided share of all NC components with independent loads, then a shuffle tree per warp and one
// over the 32 warp sums.  The scalars of the pass are finalised by the same launch (MODE 1: filter statistics,
// MODE 2: smoother statistics, MODE 0: sums only).
template <int NC, int MODE>
__global__ void __launch_bounds__(1024) k_reduce_parts_t(const real* __restrict__ part, long cnt,
                                                         real* __restrict__ out, real n, real d, int calibrate,
                                                         real* __restrict__ scal, const real* __restrict__ stop) {
  __shared__ real sh[32][NC];
  if (stop && *stop != real(0)) return;  // device-side IEKS loop ended: the scalars stay those of the final iteration
  real s[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) s[j] = 0.0;
  for (long i = threadIdx.x; i < cnt; i += 1024) {
#pragma unroll
    for (int j = 0; j < NC; ++j) s[j] += part[i * NC + j];
  }
#pragma unroll
  for (int j = 0; j < NC; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s[j] += __shfl_down_sync(0xffffffffu, s[j], o);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < NC; ++j) sh[warp][j] = s[j];
  }
  __syncthreads();
  if (warp == 0) {
    real v[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      v[j] = sh[lane][j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[j] += __shfl_down_sync(0xffffffffu, v[j], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < NC; ++j) out[j] = v[j];
      if (MODE == 1 && scal) {  // [nll, s1, s2] over n*d observations (filter.py:96-114)
        const real ssq = v[1] / n / d;
        scal[POF_S_NLL] = v[0];
        scal[POF_S_SSQ] = ssq;
        scal[POF_S_SSQ_PROPER] = v[2 < NC ? 2 : 0] / n / d;
        scal[POF_S_CSCALE] = calibrate ? sqrt(ssq) : 1.0;
      }
      if (MODE == 2 && scal) {  // [obj, #means not close]
        scal[POF_S_OBJ] = v[0];
        scal[POF_S_NOT_CLOSE] = v[1];
      }
    }
  }
}
#ifndef POF_F32
__global__ void k_finalize_seq(const real* __restrict__ sums, real n, real d, real* __restrict__ scal) {
  scal[POF_S_NLL] = -sums[0];
  scal[POF_S_SSQ] = sums[1] / n / d;
  scal[POF_S_SSQ_PROPER] = sums[2] / n / d;
  scal[POF_S_OBJ] = sums[3];
  scal[POF_S_NOT_CLOSE] = 0.0;
  scal[POF_S_CSCALE] = 1.0;
}
#endif  // !POF_F32
__global__ void k_pack_state(int D, const real* __restrict__ m, const real* __restrict__ L,
                             real* __restrict__ st) {
  for (int i = threadIdx.x; i < D + D * D; i += blockDim.x) st[i] = (i < D) ? m[i] : L[i - D];
}

// ------------------------------------------------------------------------------------------------ linearise
// one thread per step k: H_k = E1 - J E0, c_k = J y - f(y) at y = E0 m_{k+1}
__global__ void __launch_bounds__(256)
    k_linearize(int ivp_id, IvpParams P, long n, int d, int q, double scale0, double scale1,
                const real* __restrict__ means_t1, real* __restrict__ H, real* __restrict__ c) {
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int Q1 = q + 1, D = d * Q1;
  double y[4], f[4], J[16];
  for (int b = 0; b < d; ++b) y[b] = scale0 * (double)means_t1[k * D + b * Q1];
  ivp_eval(ivp_id, P, y, f, J);
  for (int a = 0; a < d; ++a) {
    double ca = -f[a];
    for (int b = 0; b < d; ++b) ca = fma(J[a * d + b], y[b], ca);
    c[k * d + a] = (real)ca;
    real* Hr = H + (k * d + a) * D;
    for (int j = 0; j < D; ++j) Hr[j] = 0.0;
    for (int b = 0; b < d; ++b) Hr[b * Q1] = (real)(-J[a * d + b] * scale0);
    Hr[a * Q1 + 1] += (real)scale1;
  }
}
// compact linearisation: per step [J_f (d x d) | c (d)] with c = J_f y - f(y); the leaf kernels rebuild H on load
__global__ void __launch_bounds__(256)
    k_linearize_compact(int ivp_id, IvpParams P, long n, int d, int q, double scale0,
                        const real* __restrict__ means_t1, real* __restrict__ Jc, const real* __restrict__ stop) {
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n || (stop && *stop != real(0))) return;
  const int Q1 = q + 1, D = d * Q1;
  double y[4], f[4], J[16];
  for (int b = 0; b < d; ++b) y[b] = scale0 * (double)means_t1[k * D + b * Q1];
  ivp_eval(ivp_id, P, y, f, J);
  real* o = Jc + k * (d * d + d);
  for (int a = 0; a < d; ++a) {
    double ca = -f[a];
    for (int b = 0; b < d; ++b) {
      ca = fma(J[a * d + b], y[b], ca);
      o[a * d + b] = (real)J[a * d + b];
    }
    o[d * d + a] = (real)ca;
  }
}
#ifndef POF_F32
// Lorenz-96 (POF_IVP_LORENZ96; the larger-state problem of BASELINE config 5, not in the reference's ivp.py):
//   f_a = (y_{a+1} - y_{a-2}) y_{a-1} - y_a + F (cyclic), 4 <= d.  One thread per (step, component): row a of the
// Jacobian has the three entries d f_a / d y_{a+1} = y_{a-1}, d f_a / d y_{a-2} = -y_{a-1}, d f_a / d y_{a-1} =
// y_{a+1} - y_{a-2} and -1 on the diagonal.  dense != 0: H (n,d,D), c (n,d); else compact [J_f | c] per step.
__global__ void __launch_bounds__(256)
    k_linearize_l96(real forcing, long n, int d, int q, double scale0, double scale1, int dense,
                    const real* __restrict__ means_t1, real* __restrict__ H, real* __restrict__ c,
                    real* __restrict__ Jc) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * d) return;
  l96_linearize_row(forcing, idx / d, (int)(idx % d), d, q, scale0, scale1, dense, means_t1, H, c, Jc);
}
#endif  // !POF_F32
// built-in problem ids and the ODE dimension each one accepts
static bool ivp_dim_ok(int ivp_id, int d) {
  static const int dims[] = {1, 2, 2, 2, 3, 3, 4, 4, 4};
#ifndef POF_F32
  if (ivp_id == POF_IVP_LORENZ96) return d >= 4 && d <= 64;
#else
  if (ivp_id == POF_IVP_LORENZ96) return false;  // large states: the fp64-only tile family
#endif
  return ivp_id >= 0 && ivp_id <= POF_IVP_HENONHEILES && dims[ivp_id] == d;
}

// ys = E0 states, with the (second) calibration multiplier of pof/solver.py:66-69 read from device memory
__global__ void __launch_bounds__(256)
    k_project(long N, int d, int q, double scale0, const real* __restrict__ mult, const real* __restrict__ means,
              const real* __restrict__ chols, real* __restrict__ ymean, real* __restrict__ ychol) {
  const int Q1 = q + 1, D = d * Q1;
  const long total = N * d * (D + 1);
  const real mu = mult ? *mult : 1.0;
  for (long idx = (long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long)gridDim.x * blockDim.x) {
    const long row = idx / (D + 1);
    const int col = (int)(idx - row * (D + 1));
    const long t = row / d;
    const int b = (int)(row - t * d);
    if (col == D) {
      ymean[row] = scale0 * means[t * D + b * Q1];
    } else if (ychol) {
      ychol[row * D + col] = mu * scale0 * chols[(t * D + b * Q1) * D + col];
    }
  }
}

// init="prior" (reference pof/initialization.py:66-89): row k >= 1 is ONE prediction of x0 = (m0, 0) over the step size
// ts[k] with the non-preconditioned model P_k F PI_k, P_k QL (transitions.py:53-77):
//   mean_k = P_k F (PI_k m0),  chol_k = tria([0, P_k QL]) = -P_k QL  (LAPACK's sign convention; the factor is
// lower triangular already), row 0 = x0.  One thread per (row, state component); the literal P F PI product is kept
// (ts[k] = 0 gives NaN exactly like the reference's 0 * inf).
__global__ void __launch_bounds__(256)
    k_prior_init(long N, int d, int q, const real* __restrict__ ts, const real* __restrict__ m0, QLParam ql,
                 real* __restrict__ means, real* __restrict__ chols) {
  const int Q1 = q + 1, D = d * Q1;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N * D) return;
  const long k = idx / D;
  const int r = (int)(idx - k * D), blk = r / Q1, i = r - blk * Q1;
  if (k == 0) {
    means[r] = m0[r];
    if (chols)
      for (int c = 0; c < D; ++c) chols[(long)r * D + c] = 0.0;
    return;
  }
  const double t = fabs((double)ts[k]);  // (evaluated in double whatever the storage type: a one-off set-up kernel)
  double fact[6];
  fact[0] = 1.0;
  for (int p = 1; p <= q; ++p) fact[p] = fact[p - 1] * p;
  // sv_j = t^(q-j+1/2) / (q-j)!,  svi_j = t^-(q-j+1/2) (q-j)!
  double acc = 0.0;
  for (int j = i; j < Q1; ++j) {
    const double svi = pow(t, -((double)(q - j) + 0.5)) * fact[q - j];
    acc = fma((double)binom(q - i, j - i), svi * (double)m0[blk * Q1 + j], acc);
  }
  const double sv = pow(t, (double)(q - i) + 0.5) / fact[q - i];
  means[k * D + r] = (real)(sv * acc);
  if (chols) {
    real* row = chols + (k * D + r) * D;
    for (int c = 0; c < D; ++c) row[c] = 0.0;
    for (int j = 0; j <= i; ++j) row[blk * Q1 + j] = (real)(-sv * (double)ql.v[i * Q1 + j]);
  }
}

#ifndef POF_F32
// 16 independent FMA chains per thread: saturates the FP64 pipe
__global__ void __launch_bounds__(256) k_dfma_peak(int iters, real* __restrict__ sink) {
  real a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const real m = 1.0 - 1e-12, c = 1e-13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
  }
  real t = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) t += a[i];
  if (t == 123.456) sink[threadIdx.x] = t;
}

#endif  // !POF_F32
// ------------------------------------------------------------------------------------------------ workspace
struct WsLayout {
  TreeLevels tl;
  long CS, L;
  int D, FE, SE, ST, NE;
  size_t o_lin, o_faggm, o_fagg, o_fin, o_sagg, o_sin, o_sx, o_kern, o_send, o_part, o_part2, o_sums, o_misc, o_flags,
      total;  // in doubles
  size_t flag_words;  // 32-bit words of the dataflow flag area (zeroed at the start of every stage)
  // hybrid filter sweep (FlowArgs in pof_launch.cuh): Kogge-Stone scan over the nodes of level ks_base, the lowest
  // level with at most KS_MAX nodes; ks_steps = 0: the tree is too shallow to gain from it.  KS_MAX measured on B200
  // (scripts/tune_tree.py, profiles/r02t_tune_tree.json): flat between 192 and 3072 -- a wider base saves levels but
  // its steps get slower (1184 concurrent combines: 12-14 us per step; 256: 8.4 us) -- 384 is best or within noise of
  // the best at every N: one combining warp per SM and step at 9472 chunks.
#if defined(POF_TUNE) && !defined(POF_F32)
  static long KS_MAX;
#else
  static constexpr long KS_MAX = 384;
#endif
  int ks_base, ks_steps;
  long ks_n;
  size_t o_ks, ks_flag_word, o_ks_s, ks_flag_word_s;  // filter scan; smoother (suffix) scan
  void build(long n, int d, int q, long chunk_len) {
    D = d * (q + 1);
    FE = 3 * D * D + 2 * D;
    SE = 2 * D * D + D;
    ST = D * D + D;
    NE = D + 2 * D * D;
    L = chunk_len < 1 ? 1 : chunk_len;
    CS = (n + L - 1) / L;
    if (CS < 1) CS = 1;
    tl.build(CS);
    size_t o = 0;
    auto take = [&](size_t cnt) {
      size_t r = o;
      o += (cnt + 31) & ~(size_t)31;
      return r;
    };
    o_lin = take((size_t)(n > 0 ? n : 1) * (d * D + d));  // linearisation of the fused iteration (compact or dense)
    o_fagg = take((size_t)tl.total * FE);
    o_faggm = take((size_t)CS * FE);
    o_fin = take((size_t)tl.total * ST);
    o_sagg = take((size_t)tl.total * SE);
    o_sin = take((size_t)tl.total * ST);
    o_sx = take((size_t)tl.total * SE);  // per node: smoothing aggregate of everything later (element-form down-sweep)
    o_kern = take((size_t)L * NE * CS);
    o_send = take((size_t)CS * ST);
    o_part = take((size_t)CS * 3);
    o_part2 = take((size_t)CS * 2);
    o_sums = take(16);
    o_misc = take((size_t)2 * ST + 64);
    ks_base = 0;
    while (ks_base < tl.nlev - 1 && tl.sz[ks_base] > KS_MAX) ++ks_base;
    ks_n = tl.sz[ks_base];
    ks_steps = 0;
    while ((1L << ks_steps) < ks_n) ++ks_steps;
    if (D > 16 || tl.nlev - 1 - ks_base < 2) ks_steps = 0;  // register-resident trees only; needs >= 2 levels to replace
    o_ks = take((size_t)ks_steps * ks_n * FE);
    ks_flag_word = 4 * (size_t)tl.total + 64;
    o_ks_s = take((size_t)ks_steps * ks_n * SE);
    ks_flag_word_s = ks_flag_word + (size_t)ks_steps * ks_n;
    flag_words = ks_flag_word_s + (size_t)ks_steps * ks_n;  // [tickets (3 x 16) | f_up | f_dn | s_up | s_dn | ks | ks_s]
    o_flags = take((flag_words * sizeof(unsigned) + sizeof(real) - 1) / sizeof(real));  // in units of the scalar type
    total = o;
  }
  unsigned* ticket(real* ws, int which) const { return (unsigned*)(ws + o_flags) + 16 * which; }
  unsigned* flags(real* ws, int which) const { return (unsigned*)(ws + o_flags) + 64 + (size_t)which * tl.total; }
};

struct ProfScope {
  pof_ctx* c;
  int idx = -1;
  cudaStream_t s;
  ProfScope(pof_ctx* ctx, int seg, cudaStream_t st) : c(ctx), s(st) {
    if (!c || !c->prof_on || c->used >= pof_ctx::MAXP) return;
    idx = c->used++;
    if (idx >= c->created) {
      cudaEventCreate(&c->ev[idx][0]);
      cudaEventCreate(&c->ev[idx][1]);
      c->created = idx + 1;
    }
    c->seg[idx] = seg;
    cudaEventRecord(c->ev[idx][0], s);
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(c->ev[idx][1], s);
  }
};

#define POF_CK(x)                     \
  do {                                \
    cudaError_t e__ = (x);            \
    if (e__ != cudaSuccess) {         \
      (void)cudaGetLastError();       \
      return (int)e__;                \
    }                                 \
  } while (0)

static int make_args(long n, int d, int q, const double* qL_host, const real* H, const real* c,
                     const WsLayout& wl, unsigned flags, LeafArgs& a) {
  if (q < 1 || q > 5) return POF_E_UNSUPPORTED_DQ;
  a.n = n;
  a.L = wl.L;
  a.CS = wl.CS;
  a.H = H;
  a.c = c;
  a.Jc = nullptr;
  a.R = nullptr;
  a.F = nullptr;
  a.QLd = nullptr;
  a.tile_reg = (flags & POF_F_TILE_SMEM_QR) ? 0 : 1;
  a.no_tma = (flags & POF_F_SMOOTH_TMA) ? 0 : 1;
  a.stop = nullptr;
  a.d = d;
  a.q = q;
  a.s0 = a.s1 = 0.0;
  for (int i = 0; i < 36; ++i) a.ql.v[i] = 0.0;
  for (int i = 0; i < (q + 1) * (q + 1); ++i) a.ql.v[i] = qL_host[i];
  return 0;
}

#if defined(POF_TUNE) && !defined(POF_F32)
long WsLayout::KS_MAX = 384;
int g_poll_ns = 64;
extern "C" void pof_tune_ks_max(long v) { WsLayout::KS_MAX = v; }
extern "C" void pof_tune_poll_ns(int v) { g_poll_ns = v; }
// out: [flag area offset in bytes, nodes in total, levels, ks_base, ks_steps, ks_n, ks_flag_word, sz[0..levels)]
extern "C" void pof_tune_flag_layout(long n, int d, int q, long chunk_len, long* out) {
  WsLayout wl;
  wl.build(n, d, q, chunk_len);
  out[0] = (long)(wl.o_flags * sizeof(real));
  out[1] = wl.tl.total;
  out[2] = wl.tl.nlev;
  out[3] = wl.ks_base;
  out[4] = wl.ks_steps;
  out[5] = wl.ks_n;
  out[6] = (long)wl.ks_flag_word;
  for (int l = 0; l < wl.tl.nlev; ++l) out[7 + l] = wl.tl.sz[l];
}
#endif
// ---- dataflow sweeps (FlowArgs, pof_launch.cuh)
static void flow_begin(FlowArgs& fa, const WsLayout& wl, real* agg, real* st, unsigned* f_up, unsigned* f_dn,
                       unsigned* ticket) {
  fa.nlev = wl.tl.nlev;
  for (int l = 0; l < FlowArgs::MAXL; ++l) {
    fa.off[l] = l < wl.tl.nlev ? wl.tl.off[l] : 0;
    fa.sz[l] = l < wl.tl.nlev ? wl.tl.sz[l] : 0;
  }
  fa.nseg = 0;
  fa.up_lo = 1;
  fa.up_hi = 0;
  fa.agg = agg;
  fa.st = st;
  fa.sx = nullptr;
  fa.root_m = fa.root_L = nullptr;
  fa.flag_up = f_up;
  fa.flag_dn = f_dn;
  fa.ticket = ticket;
  fa.stop = nullptr;
  fa.ks_base = fa.ks_steps = 0;
  fa.ks_n = 0;
  fa.ks = nullptr;
  fa.flag_ks = nullptr;
  fa.ks_wait = 0;
#if defined(POF_TUNE) && !defined(POF_F32)
  fa.poll_ns = g_poll_ns;
#else
  fa.poll_ns = 64;
#endif
}
// hybrid filter sweep, first half: up-sweep to level ks_base, Kogge-Stone scan over its nodes (the last node's final
// element is the aggregate of the whole sequence: the time-sharded form's carry, ks_total())
static void flow_seg(FlowArgs& fa, int kind, int level, long count) {
  fa.seg_kind[fa.nseg] = kind;
  fa.seg_level[fa.nseg] = level;
  fa.seg_count[fa.nseg] = count;
  ++fa.nseg;
}
static bool use_hybrid(const WsLayout& wl, unsigned flags) { return wl.ks_steps > 0 && !(flags & POF_F_TREE_UPDOWN); }
static void flow_hybrid_fields(FlowArgs& fa, const WsLayout& wl, real* ws, int ks_wait) {
  fa.ks_base = wl.ks_base;
  fa.ks_steps = wl.ks_steps;
  fa.ks_n = wl.ks_n;
  fa.ks = ws + wl.o_ks;
  fa.flag_ks = (unsigned*)(ws + wl.o_flags) + wl.ks_flag_word;
  fa.ks_wait = ks_wait;
}
static void flow_hybrid_up(FlowArgs& fa, const WsLayout& wl, real* ws) {
  flow_hybrid_fields(fa, wl, ws, 1);
  if (wl.ks_base >= 1) {
    fa.up_lo = 1;
    fa.up_hi = wl.ks_base;
    for (int l = 1; l <= wl.ks_base; ++l) flow_seg(fa, FlowArgs::UP, l, wl.tl.sz[l]);
  }
  for (int st = 1; st <= wl.ks_steps; ++st) flow_seg(fa, FlowArgs::KS, st, wl.ks_n - (1L << (st - 1)));
}
static const real* ks_total(const WsLayout& wl, const real* ws) {  // ks_steps >= 1: the last node's level is ks_steps
  return ws + wl.o_ks + ((size_t)(wl.ks_steps - 1) * wl.ks_n + (wl.ks_n - 1)) * wl.FE;
}
// second half: states of level ks_base from the root state, down-sweep from there.  ks_wait = 0: the Kogge-Stone
// elements were completed by an earlier launch.
static void flow_hybrid_down(FlowArgs& fa, const WsLayout& wl, real* ws, const real* root_m, const real* root_L,
                             int ks_wait) {
  flow_hybrid_fields(fa, wl, ws, ks_wait);
  fa.root_m = root_m;
  fa.root_L = root_L;
  flow_seg(fa, FlowArgs::ROOT, wl.tl.nlev - 1, 1);
  flow_seg(fa, FlowArgs::KS_APPLY, wl.ks_base, wl.ks_n);
  for (int l = wl.ks_base; l >= 1; --l) flow_seg(fa, FlowArgs::DOWN, l, wl.tl.sz[l]);
}
// up-sweep: build levels 1 .. top from their children
static void flow_up(FlowArgs& fa, const WsLayout& wl, int top) {
  if (top < 1) return;
  fa.up_lo = 1;
  fa.up_hi = top;
  for (int l = 1; l <= top; ++l) {
    fa.seg_kind[fa.nseg] = FlowArgs::UP;
    fa.seg_level[fa.nseg] = l;
    fa.seg_count[fa.nseg] = wl.tl.sz[l];
    ++fa.nseg;
  }
}
// down-sweep from the root state (root_m, root_L): states of all nodes, level nlev-1 .. 0
static void flow_down(FlowArgs& fa, const WsLayout& wl, const real* root_m, const real* root_L) {
  fa.root_m = root_m;
  fa.root_L = root_L;
  fa.seg_kind[fa.nseg] = FlowArgs::ROOT;
  fa.seg_level[fa.nseg] = wl.tl.nlev - 1;
  fa.seg_count[fa.nseg] = 1;
  ++fa.nseg;
  for (int l = wl.tl.nlev - 1; l >= 1; --l) {
    fa.seg_kind[fa.nseg] = FlowArgs::DOWN;
    fa.seg_level[fa.nseg] = l;
    fa.seg_count[fa.nseg] = wl.tl.sz[l];
    ++fa.nseg;
  }
}
// smoother only: element-form down-sweep -- per node the aggregate of all LATER nodes (identity at the root)
static void flow_down_elem(FlowArgs& fa, const WsLayout& wl, real* sx) {
  fa.sx = sx;
  fa.root_m = fa.root_L = nullptr;
  fa.seg_kind[fa.nseg] = FlowArgs::ROOT;
  fa.seg_level[fa.nseg] = wl.tl.nlev - 1;
  fa.seg_count[fa.nseg] = 1;
  ++fa.nseg;
  for (int l = wl.tl.nlev - 1; l >= 1; --l) {
    fa.seg_kind[fa.nseg] = FlowArgs::DOWN_E;
    fa.seg_level[fa.nseg] = l;
    fa.seg_count[fa.nseg] = wl.tl.sz[l];
    ++fa.nseg;
  }
}
// smoother, element-form suffix scan as a hybrid sweep: up-sweep to level ks_base, Kogge-Stone SUFFIX scan over its
// nodes, "everything later" aggregates of that level, element-form down-sweep from there
static void flow_hybrid_suffix(FlowArgs& fa, const WsLayout& wl, real* ws) {
  flow_hybrid_fields(fa, wl, ws, 1);
  fa.ks = ws + wl.o_ks_s;
  fa.flag_ks = (unsigned*)(ws + wl.o_flags) + wl.ks_flag_word_s;
  fa.sx = ws + wl.o_sx;
  fa.root_m = fa.root_L = nullptr;
  if (wl.ks_base >= 1) {
    fa.up_lo = 1;
    fa.up_hi = wl.ks_base;
    for (int l = 1; l <= wl.ks_base; ++l) flow_seg(fa, FlowArgs::UP, l, wl.tl.sz[l]);
  }
  for (int st = 1; st <= wl.ks_steps; ++st) flow_seg(fa, FlowArgs::KS, st, wl.ks_n - (1L << (st - 1)));
  flow_seg(fa, FlowArgs::KS_APPLY, wl.ks_base, wl.ks_n);
  for (int l = wl.ks_base; l >= 1; --l) flow_seg(fa, FlowArgs::DOWN_E, l, wl.tl.sz[l]);
}
// Element-form suffix scan of the smoother (its whole tree work runs concurrently with the filter scan, one state
// combine per chunk remains afterwards) pays when the filter scan is long enough to hide it: 2 x log2(#chunks)
// general smoothing combines against log2(#chunks) cheaper state-form ones after the scan.  Measured on B200 (FHN,
// D = 8): break-even at ~40 steps per chunk (N ~ 2^18.5); below that the state-form down-sweep stays.
// With the hybrid sweep (one GPU) the element-form scan has ~half the dependent steps: it then also pays for shorter
// chunks as long as there are enough of them (>= 1024: below that the state-form sweep after the scan is as short).
static bool hybrid_suffix(const WsLayout& wl, unsigned flags, bool sharded) {
  return !sharded && use_hybrid(wl, flags) && (wl.L >= 48 || wl.CS >= 1024);
}
static bool elem_suffix(const WsLayout& wl, unsigned flags, bool sharded) {
  return wl.L >= 48 || hybrid_suffix(wl, flags, sharded);
}
static int zero_flags(cudaStream_t s, const WsLayout& wl, real* ws) {
  POF_CK(cudaMemsetAsync(ws + wl.o_flags, 0, wl.flag_words * sizeof(unsigned), s));
  return 0;
}

enum { TK_FILTER = 0, TK_SUP = 1, TK_SDOWN = 2 };
enum { FL_FUP = 0, FL_FDN = 1, FL_SUP = 2, FL_SDN = 3 };

// stage A: fold (+ in the time-sharded form, need_root: the filter up-sweep whose root is the shard's carry element;
// on one GPU the up-sweep runs in the same dataflow launch as the down-sweep, in stage B)
static int stage_a(cudaStream_t s, pof_ctx* ctx, unsigned flags, const LeafLaunch* ll, const LeafArgs& a,
                   const WsLayout& wl, real* ws, bool need_root) {
  real* fagg = ws + wl.o_fagg;
  {
    ProfScope ps(ctx, POF_SEG_FOLD, s);
    POF_CK(ll->fold(s, a, fagg, ws + wl.o_faggm));
  }
  if (!need_root) return 0;
  const TreeLaunch* tl = tree_for(ll, wl.D, flags);
  ProfScope ps(ctx, POF_SEG_FUP, s);
  if (tl && !(flags & POF_F_TREE_PER_LEVEL)) {
    FlowArgs fa;
    flow_begin(fa, wl, fagg, ws + wl.o_fin, wl.flags(ws, FL_FUP), wl.flags(ws, FL_FDN), wl.ticket(ws, TK_FILTER));
    fa.stop = a.stop;
    if (use_hybrid(wl, flags)) flow_hybrid_up(fa, wl, ws);
    else flow_up(fa, wl, wl.tl.nlev - 1);
    if (fa.nseg) POF_CK(tl->fflow(s, fa));
    return 0;
  }
  for (int l = 0; l + 1 < wl.tl.nlev; ++l) {
    const long np = wl.tl.sz[l + 1];
    if (tl)
      POF_CK(tl->fup(s, fagg + wl.tl.off[l] * wl.FE, wl.tl.sz[l], nullptr, fagg + wl.tl.off[l + 1] * wl.FE, np));
    else
      POF_CK(tile_fup(s, wl.D, fagg + wl.tl.off[l] * wl.FE, wl.tl.sz[l], fagg + wl.tl.off[l + 1] * wl.FE, np));
  }
  return (int)cudaGetLastError();
}

// stage B: filter tree (single GPU: up-sweep + down-sweep; sharded: down-sweep only) from the root's incoming state
// (root_m, root_L), then the filter scan on `s` while the chunk-level smoothing elements and the smoother's up-sweep
// run on the context's side stream (their inputs only depend on the fold and the filter tree), join, filter scalars.
static int stage_b(cudaStream_t s, pof_ctx* ctx, unsigned flags, const LeafLaunch* ll, const LeafArgs& a,
                   const WsLayout& wl, real* ws, const real* root_m, const real* root_L, real* fmeans,
                   real* fchols, bool need_root, real n_obs_total, int calibrate, real* scalars) {
  real* fagg = ws + wl.o_fagg;
  real* fin = ws + wl.o_fin;
  real* sagg = ws + wl.o_sagg;
  const TreeLaunch* tl = tree_for(ll, wl.D, flags);
  const bool per_level = !tl || (flags & POF_F_TREE_PER_LEVEL);
  const int up_top = wl.tl.nlev - 1 - (need_root ? 0 : 1);  // the root element is only consumed by the sharded form
  {
    ProfScope ps(ctx, POF_SEG_FTREE, s);
    if (!per_level) {
      FlowArgs fa;
      flow_begin(fa, wl, fagg, fin, wl.flags(ws, FL_FUP), wl.flags(ws, FL_FDN), wl.ticket(ws, TK_FILTER));
      fa.stop = a.stop;
      if (use_hybrid(wl, flags)) {
        if (!need_root) flow_hybrid_up(fa, wl, ws);
        flow_hybrid_down(fa, wl, ws, root_m, root_L, need_root ? 0 : 1);
      } else {
        if (!need_root) flow_up(fa, wl, up_top);
        flow_down(fa, wl, root_m, root_L);
      }
      POF_CK(tl->fflow(s, fa));
    } else {
      if (!need_root)
        for (int l = 0; l < up_top; ++l) {
          const long np = wl.tl.sz[l + 1];
          if (tl)
            POF_CK(tl->fup(s, fagg + wl.tl.off[l] * wl.FE, wl.tl.sz[l], nullptr, fagg + wl.tl.off[l + 1] * wl.FE, np));
          else
            POF_CK(tile_fup(s, wl.D, fagg + wl.tl.off[l] * wl.FE, wl.tl.sz[l], fagg + wl.tl.off[l + 1] * wl.FE, np));
        }
      k_pack_state<<<1, 128, 0, s>>>(wl.D, root_m, root_L, fin + wl.tl.off[wl.tl.nlev - 1] * wl.ST);
      for (int l = wl.tl.nlev - 1; l >= 1; --l) {
        const long np = wl.tl.sz[l];
        if (tl)
          POF_CK(tl->fdown(s, fin + wl.tl.off[l] * wl.ST, wl.tl.sz[l - 1], fagg + wl.tl.off[l - 1] * wl.FE,
                           fin + wl.tl.off[l - 1] * wl.ST, np));
        else
          POF_CK(tile_fdown(s, wl.D, fin + wl.tl.off[l] * wl.ST, np, fagg + wl.tl.off[l - 1] * wl.FE, wl.tl.sz[l - 1],
                            fin + wl.tl.off[l - 1] * wl.ST));
      }
    }
  }
  POF_CK(cudaGetLastError());
  // chunk-level smoothing elements straight from (incoming state, filtering element before its last update), then
  // the smoother's up-sweep: on the side stream if the caller gave a context, else in line after the scan
  auto smoother_up = [&](cudaStream_t st) -> int {
    ProfScope ps(ctx, POF_SEG_SUP, st);
    if (tl)
      POF_CK(tl->chunkk(st, fin, wl.CS, ws + wl.o_faggm, sagg, wl.CS));
    else
      POF_CK(tile_chunkk(st, wl.D, fin, ws + wl.o_faggm, sagg, wl.CS));
    if (!per_level) {
      // The smoothing elements do not depend on the terminal state, so the WHOLE suffix scan runs here in ELEMENT
      // form -- up-sweep, then down-sweep giving every chunk the aggregate of all later chunks -- concurrently with
      // the filter scan; once the terminal state exists, ONE state-form combine per chunk remains (stage C).
      FlowArgs fa;
      flow_begin(fa, wl, sagg, ws + wl.o_sin, wl.flags(ws, FL_SUP), wl.flags(ws, FL_SDN), wl.ticket(ws, TK_SUP));
      fa.stop = a.stop;
      if (hybrid_suffix(wl, flags, need_root)) {
        flow_hybrid_suffix(fa, wl, ws);
      } else {
        flow_up(fa, wl, up_top);
        if (elem_suffix(wl, flags, need_root)) flow_down_elem(fa, wl, ws + wl.o_sx);
      }
      if (fa.nseg) POF_CK(tl->sflow(st, fa));
      return 0;
    }
    for (int l = 0; l < up_top; ++l) {
      const long np = wl.tl.sz[l + 1];
      if (tl)
        POF_CK(tl->sup(st, sagg + wl.tl.off[l] * wl.SE, wl.tl.sz[l], nullptr, sagg + wl.tl.off[l + 1] * wl.SE, np));
      else
        POF_CK(tile_sup(st, wl.D, sagg + wl.tl.off[l] * wl.SE, wl.tl.sz[l], sagg + wl.tl.off[l + 1] * wl.SE, np));
    }
    return (int)cudaGetLastError();
  };
  const bool side = ctx && ctx->s2 && tl;
  if (side) {
    POF_CK(cudaEventRecord(ctx->fork, s));
    POF_CK(cudaStreamWaitEvent(ctx->s2, ctx->fork, 0));
    if (int rc = smoother_up(ctx->s2)) return rc;
    POF_CK(cudaEventRecord(ctx->join, ctx->s2));
  }
  {
    ProfScope ps(ctx, POF_SEG_SCAN, s);
    POF_CK(ll->scan(s, a, fin, ws + wl.o_kern, nullptr, ws + wl.o_send, ws + wl.o_part, fmeans, fchols));
  }
  if (side)
    POF_CK(cudaStreamWaitEvent(s, ctx->join, 0));
  else if (int rc = smoother_up(s))
    return rc;
  k_reduce_parts_t<3, 1><<<1, 1024, 0, s>>>(ws + wl.o_part, wl.CS, ws + wl.o_sums, n_obs_total, (real)a.d, calibrate,
                                            scalars, a.stop);
  return (int)cudaGetLastError();
}

// stage C: smoother down-sweep from the seed (root_m, root_L) + smoother scan + smoother scalars
static int stage_c(cudaStream_t s, pof_ctx* ctx, unsigned flags, const LeafLaunch* ll, const LeafArgs& a,
                   const WsLayout& wl, real* ws, const real* root_m, const real* root_L, int emit_t0,
                   const real* cscale, real* means, real* chols, real* scalars, bool sharded) {
  real* sagg = ws + wl.o_sagg;
  real* sin_ = ws + wl.o_sin;
  const TreeLaunch* tl = tree_for(ll, wl.D, flags);
  const bool per_level = !tl || (flags & POF_F_TREE_PER_LEVEL);
  {
    ProfScope ps(ctx, POF_SEG_SDOWN, s);
    if (!per_level) {
      if (elem_suffix(wl, flags, sharded)) {
        // seeds of all chunks at once: (terminal state) combined with the chunk's "everything later" aggregate
        POF_CK(tl->sseed(s, root_m, 1, ws + wl.o_sx, sin_, wl.CS));
      } else {
        FlowArgs fa;
        flow_begin(fa, wl, sagg, sin_, wl.flags(ws, FL_SUP), wl.flags(ws, FL_SDN), wl.ticket(ws, TK_SDOWN));
        fa.stop = a.stop;
        flow_down(fa, wl, root_m, root_L);
        POF_CK(tl->sflow(s, fa));
      }
    } else {
      k_pack_state<<<1, 128, 0, s>>>(wl.D, root_m, root_L, sin_ + wl.tl.off[wl.tl.nlev - 1] * wl.ST);
      for (int l = wl.tl.nlev - 1; l >= 1; --l) {
        const long np = wl.tl.sz[l];
        if (tl)
          POF_CK(tl->sdown(s, sin_ + wl.tl.off[l] * wl.ST, wl.tl.sz[l - 1], sagg + wl.tl.off[l - 1] * wl.SE,
                           sin_ + wl.tl.off[l - 1] * wl.ST, np));
        else
          POF_CK(tile_sdown(s, wl.D, sin_ + wl.tl.off[l] * wl.ST, np, sagg + wl.tl.off[l - 1] * wl.SE, wl.tl.sz[l - 1],
                            sin_ + wl.tl.off[l - 1] * wl.ST));
      }
    }
  }
  POF_CK(cudaGetLastError());
  {
    ProfScope ps(ctx, POF_SEG_SMOOTH, s);
    POF_CK(ll->smooth(s, a, sin_, ws + wl.o_kern, emit_t0, cscale, means, chols, ws + wl.o_part2));
  }
  k_reduce_parts_t<2, 2><<<1, 1024, 0, s>>>(ws + wl.o_part2, wl.CS, ws + wl.o_sums + 8, 0.0, 0.0, 0, scalars, a.stop);
  return (int)cudaGetLastError();
}

static int run_pass(cudaStream_t s, pof_ctx* ctx, unsigned flags, const LeafLaunch* ll, const LeafArgs& a,
                    const WsLayout& wl, real* ws, int64_t N, const real* x0_mean, const real* x0_chol,
                    real* means, real* chols, real* fmeans, real* fchols, int calibrate, real* scalars) {
  if (int rc = zero_flags(s, wl, ws)) return rc;
  if (int rc = stage_a(s, ctx, flags, ll, a, wl, ws, false)) return rc;
  if (fmeans) {
    POF_CK(cudaMemcpyAsync(fmeans, x0_mean, wl.D * sizeof(real), cudaMemcpyDeviceToDevice, s));
    POF_CK(cudaMemcpyAsync(fchols, x0_chol, wl.D * wl.D * sizeof(real), cudaMemcpyDeviceToDevice, s));
  }
  if (int rc = stage_b(s, ctx, flags, ll, a, wl, ws, x0_mean, x0_chol, fmeans, fchols, false, (real)(N - 1), calibrate,
                       scalars))
    return rc;
  // terminal smoothing state = filtered state at the last time point
  const real* term = ws + wl.o_send + (wl.CS - 1) * wl.ST;
  return stage_c(s, ctx, flags, ll, a, wl, ws, term, term + wl.D, 1, scalars + POF_S_CSCALE, means, chols, scalars,
                 false);
}

static int fill_params(int ivp_id, const double* params_host, int nparams, int d, IvpParams& P) {
  if (ivp_id < 0 || ivp_id > POF_IVP_LORENZ96) return POF_E_IVP;
  if (nparams > 8 || !ivp_dim_ok(ivp_id, d)) return POF_E_ARG;
  for (int i = 0; i < 8; ++i) P.p[i] = (i < nparams) ? params_host[i] : 0.0;
  return 0;
}

}  // namespace POF_NS

using namespace POF_NS;

extern "C" {

#ifndef POF_F32
int pof_supported(int d, int q) { return leaf_launch(d, q, 0) != nullptr ? 1 : 0; }
int pof_supported_tile(int d, int q) { return tile_supported(d, q) ? 1 : 0; }

// ---- context
int pof_ctx_create(pof_ctx_t** out) {
  if (!out) return POF_E_ARG;
  pof_ctx* c = new pof_ctx();
  cudaError_t e = cudaGetDevice(&c->dev);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->s2, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->join, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    delete c;
    return (int)e;
  }
  *out = c;
  return 0;
}
void pof_ctx_destroy(pof_ctx_t* c) {
  if (!c) return;
  for (int i = 0; i < c->created; ++i) {
    cudaEventDestroy(c->ev[i][0]);
    cudaEventDestroy(c->ev[i][1]);
  }
  if (c->fork) cudaEventDestroy(c->fork);
  if (c->join) cudaEventDestroy(c->join);
  if (c->s2) cudaStreamDestroy(c->s2);
  delete c;
}
void pof_ctx_profile_enable(pof_ctx_t* c, int on) {
  if (!c) return;
  c->prof_on = on != 0;
  c->used = 0;
  for (int i = 0; i < POF_SEG_COUNT; ++i) {
    c->acc[i] = 0.0;
    c->cnt[i] = 0;
  }
}
// synchronises the device; ms_out[POF_SEG_COUNT] = accumulated milliseconds per segment kind, count_out = #segments
int pof_ctx_profile_read(pof_ctx_t* c, double* ms_out, int64_t* count_out) {
  if (!c) return POF_E_ARG;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return (int)e;
  for (int i = 0; i < c->used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev[i][0], c->ev[i][1]) == cudaSuccess) {
      c->acc[c->seg[i]] += ms;
      c->cnt[c->seg[i]] += 1;
    }
  }
  c->used = 0;
  for (int i = 0; i < POF_SEG_COUNT; ++i) {
    ms_out[i] = c->acc[i];
    count_out[i] = c->cnt[i];
  }
  return 0;
}

// kernels launched by one pof_linear_filtsmooth_f64 call (memset / memcpy nodes not counted)
int64_t pof_launches_per_pass(int64_t N, int d, int q, int64_t chunk_len, uint32_t flags) {
  WsLayout wl;
  wl.build(N - 1, d, q, chunk_len);
  const LeafLaunch* ll = leaf_launch(d, q, flags);
  if (!ll) return 0;
  const TreeLaunch* tl = tree_for(ll, wl.D, flags);
  const int64_t leaf = 3, chunkk = 1, reduce = 2;
  if (tl && !(flags & POF_F_TREE_PER_LEVEL)) {
    // smoother: element-form suffix scan (1) + chunk seeds (1), or up-sweep (1 if it has a level to build) + down-sweep
    const int64_t sm = elem_suffix(wl, flags, false) ? 2 : ((wl.tl.nlev >= 3 ? 1 : 0) + 1);
    return leaf + chunkk + reduce + 1 /*filter tree*/ + sm;
  }
  const int up_total = wl.tl.nlev >= 2 ? wl.tl.nlev - 2 : 0;  // the root combine is skipped on one GPU
  const int down_total = wl.tl.nlev - 1;
  return leaf + chunkk + reduce + 2 /*pack*/ + 2 * (int64_t)(up_total + down_total);
}

// FP64 FMA throughput of this device (TFLOP/s), measured with a register-resident DFMA loop: the roofline
// denominator for the FP64-bound kernels (MEASURED_PEAKS.json holds no FP64 number)
int pof_measure_dfma_tflops(pof_stream_t s_, double* tflops_out) {
  cudaStream_t s = (cudaStream_t)s_;
  int dev = 0, sms = 0;
  POF_CK(cudaGetDevice(&dev));
  POF_CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  real* sink = nullptr;
  POF_CK(cudaMalloc(&sink, sizeof(real) * 1024));
  const int iters = 1 << 14, blocks = sms * 8, threads = 256;
  cudaEvent_t e0, e1;
  POF_CK(cudaEventCreate(&e0));
  POF_CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    POF_CK(cudaEventRecord(e0, s));
    k_dfma_peak<<<blocks, threads, 0, s>>>(iters, sink);
    POF_CK(cudaEventRecord(e1, s));
    POF_CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    POF_CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(sink);
  const real flops = 2.0 * 16.0 * (real)iters * (real)blocks * (real)threads;
  *tflops_out = flops / (best * 1e-3) / 1e12;
  return 0;
}

int64_t pof_default_chunk_len(int64_t N, int d, int q, int sm_count, uint32_t flags) {
  if (sm_count <= 0) sm_count = 148;
  const LeafLaunch* ll = leaf_launch(d, q, flags);
  const int64_t n = N - 1;
  // lane2: ~8 resident warps per SM, chunks_per_warp chunks per warp; tile family: one chunk per resident CTA
  const int64_t target = (!ll || ll->is_tile) ? (int64_t)sm_count * tile_ctas_per_sm(d, q)
                                               : (int64_t)sm_count * 8 * ll->chunks_per_warp;
  int64_t L = (n + target - 1) / target;
  if (L < 4) L = 4;
  return L;
}

size_t pof_workspace_bytes(int64_t N, int d, int q, int64_t chunk_len) {
  WsLayout wl;
  wl.build(N - 1, d, q, chunk_len);
  return wl.total * sizeof(real);
}

#else
size_t pof_workspace_bytes_f32(int64_t N, int d, int q, int64_t chunk_len) {
  WsLayout wl;
  wl.build(N - 1, d, q, chunk_len);
  return wl.total * sizeof(real);
}
#endif  // !POF_F32

int POF_SUFFIX(pof_filter_combine)(pof_stream_t s, int64_t n, int D, const real* e1, const real* e2, real* out,
                           uint32_t flags) {
  if (n <= 0) return 0;
  if (!(flags & POF_F_FAMILY_TILE))
    if (const TreeLaunch* tl = tree_launch(D)) return (int)tl->fcomb((cudaStream_t)s, e1, n, e2, out, n);
  if (!tile_tree_supported(D)) return POF_E_UNSUPPORTED_DQ;
  return (int)tile_fcomb((cudaStream_t)s, D, n, e1, e2, out);
}
int POF_SUFFIX(pof_smooth_combine)(pof_stream_t s, int64_t n, int D, const real* e1, const real* e2, real* out,
                           uint32_t flags) {
  if (n <= 0) return 0;
  if (!(flags & POF_F_FAMILY_TILE))
    if (const TreeLaunch* tl = tree_launch(D)) return (int)tl->scomb((cudaStream_t)s, e1, n, e2, out, n);
  if (!tile_tree_supported(D)) return POF_E_UNSUPPORTED_DQ;
  return (int)tile_scomb((cudaStream_t)s, D, n, e1, e2, out);
}

int POF_SUFFIX(pof_linearize_ivp)(pof_stream_t s, int ivp_id, const double* params_host, int nparams, int64_t n, int d, int q,
                          double scale0, double scale1, const real* means_t1, real* H, real* c) {
  IvpParams P;
  if (int rc = fill_params(ivp_id, params_host, nparams, d, P)) return rc;
  if (n <= 0) return 0;
#ifndef POF_F32
  if (ivp_id == POF_IVP_LORENZ96) {
    k_linearize_l96<<<(unsigned)((n * d + 255) / 256), 256, 0, (cudaStream_t)s>>>(P.p[0], n, d, q, scale0, scale1, 1,
                                                                                  means_t1, H, c, nullptr);
    return (int)cudaGetLastError();
  }
#endif
  k_linearize<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(ivp_id, P, n, d, q, scale0, scale1, means_t1,
                                                                        H, c);
  return (int)cudaGetLastError();
}

int POF_SUFFIX(pof_linearize_ivp_compact)(pof_stream_t s, int ivp_id, const double* params_host, int nparams, int64_t n, int d,
                                  int q, double scale0, const real* means_t1, real* Jc) {
  IvpParams P;
  if (int rc = fill_params(ivp_id, params_host, nparams, d, P)) return rc;
  if (n <= 0) return 0;
#ifndef POF_F32
  if (ivp_id == POF_IVP_LORENZ96) {
    k_linearize_l96<<<(unsigned)((n * d + 255) / 256), 256, 0, (cudaStream_t)s>>>(P.p[0], n, d, q, scale0, 0.0, 0,
                                                                                  means_t1, nullptr, nullptr, Jc);
    return (int)cudaGetLastError();
  }
#endif
  k_linearize_compact<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(ivp_id, P, n, d, q, scale0, means_t1,
                                                                                Jc, nullptr);
  return (int)cudaGetLastError();
}

int POF_SUFFIX(pof_linear_filtsmooth)(pof_stream_t s_, pof_ctx_t* ctx, uint32_t flags, int64_t N, int d, int q,
                              int64_t chunk_len, const double* qL_host, const real* x0_mean, const real* x0_chol,
                              const real* H, const real* c, real* means, real* chols, real* fmeans,
                              real* fchols, int calibrate, real* scalars, void* ws_, size_t ws_bytes) {
  cudaStream_t s = (cudaStream_t)s_;
  if (N < 2) return POF_E_ARG;
  const LeafLaunch* ll = leaf_launch(d, q, flags);
  if (!ll) return POF_E_UNSUPPORTED_DQ;
  WsLayout wl;
  wl.build(N - 1, d, q, chunk_len);
  if (ws_bytes < wl.total * sizeof(real)) return POF_E_WORKSPACE;
  LeafArgs a;
  if (int rc = make_args(N - 1, d, q, qL_host, H, c, wl, flags, a)) return rc;
  return run_pass(s, ctx, flags, ll, a, wl, (real*)ws_, N, x0_mean, x0_chol, means, chols, fmeans, fchols, calibrate,
                  scalars);
}

#ifndef POF_F32
// general linear-Gaussian model: always the tile family (the only leaves that carry the D-column posterior factor and
// read dense per-step transition models)
int POF_SUFFIX(pof_linear_filtsmooth_general)(pof_stream_t s_, pof_ctx_t* ctx, uint32_t flags, int64_t N, int d, int q,
                                      int64_t chunk_len, const double* qL_host, const real* F, const real* QL,
                                      const real* x0_mean, const real* x0_chol, const real* H, const real* c,
                                      const real* cholR, real* means, real* chols, real* fmeans, real* fchols,
                                      int calibrate, real* scalars, void* ws_, size_t ws_bytes) {
  cudaStream_t s = (cudaStream_t)s_;
  if (N < 2) return POF_E_ARG;
  if ((F == nullptr) != (QL == nullptr)) return POF_E_ARG;
  if (!F && !qL_host) return POF_E_ARG;
  if (!tile_supported(d, q)) return POF_E_UNSUPPORTED_DQ;
  flags |= POF_F_FAMILY_TILE;
  const LeafLaunch* ll = tile_leaf_launch();
  WsLayout wl;
  wl.build(N - 1, d, q, chunk_len);
  if (ws_bytes < wl.total * sizeof(real)) return POF_E_WORKSPACE;
  LeafArgs a;
  double ql_dummy[36] = {0.0};
  if (int rc = make_args(N - 1, d, q, qL_host ? qL_host : ql_dummy, H, c, wl, flags, a)) return rc;
  a.R = cholR;
  a.F = F;
  a.QLd = QL;
  return run_pass(s, ctx, flags, ll, a, wl, (real*)ws_, N, x0_mean, x0_chol, means, chols, fmeans, fchols, calibrate,
                  scalars);
}

#endif  // !POF_F32

// loop_state (device, 8 values; null = a single iteration): [0] stop flag, [1] iterations done, [2] obj of the last
// iteration, [3] nll of the last iteration.  With it every kernel returns at once when the flag is set, and k_crit
// evaluates the reference's stopping rule (convergence_criteria.py:4-13, solver.py:36-45) on the device.
// In a loop graph (pof_ieks_loop_create) it also sets the WHILE node's condition.
static __global__ void k_crit(const real* __restrict__ scal, real* __restrict__ ls, double maxiters,
                              cudaGraphConditionalHandle cond, int has_cond) {
  if (ls[0] != real(0)) {
    if (has_cond) cudaGraphSetConditional(cond, 0u);
    return;
  }
  const double k = (double)ls[1] + 1.0;
  const double obj = (double)scal[POF_S_OBJ], nll = (double)scal[POF_S_NLL], bad = (double)scal[POF_S_NOT_CLOSE];
  const double obj_old = (double)ls[2];
  const bool isnan_ = (obj != obj) || (nll != nll);
  // numpy / jax isclose(a = obj_old, b = obj): |a - b| <= atol + rtol |b|, rtol 1e-6, atol 1e-9
  const bool obj_conv = !isnan_ && fabs(obj_old - obj) <= 1e-9 + 1e-6 * fabs(obj);
  const bool conv = isnan_ || obj_conv || bad == 0.0;
  ls[1] = (real)k;
  ls[2] = (real)obj;
  ls[3] = (real)nll;
  const bool stop = conv || !(k <= maxiters);
  ls[0] = stop ? real(1) : real(0);
  if (has_cond) cudaGraphSetConditional(cond, stop ? 0u : 1u);
}

static int ieks_iteration_impl(cudaStream_t s, pof_ctx_t* ctx, uint32_t flags, int ivp_id, const double* params_host,
                               int nparams, int64_t N, int d, int q, int64_t chunk_len, const double* qL_host,
                               double scale0, double scale1, const real* x0_mean, const real* x0_chol, real* means,
                               real* chols, int calibrate, real* scalars, real* loop_state, double maxiters, void* ws_,
                               size_t ws_bytes, const cudaGraphConditionalHandle* cond = nullptr) {
  if (N < 2) return POF_E_ARG;
  IvpParams P;
  if (int rc = fill_params(ivp_id, params_host, nparams, d, P)) return rc;
  const LeafLaunch* ll = leaf_launch(d, q, flags);
  if (!ll) return POF_E_UNSUPPORTED_DQ;
  if (loop_state && tree_for(ll, d * (q + 1), flags) == nullptr) return POF_E_UNSUPPORTED_DQ;  // register-resident family
  WsLayout wl;
  wl.build(N - 1, d, q, chunk_len);
  if (ws_bytes < wl.total * sizeof(real)) return POF_E_WORKSPACE;
  real* ws = (real*)ws_;
  const long n = N - 1;
  const int D = wl.D;
  real* lin = ws + wl.o_lin;
  // compact linearisation [J_f | c] per step; both kernel families rebuild H = E1 - J_f E0 on load
#ifndef POF_F32
  if (ivp_id == POF_IVP_LORENZ96)
    k_linearize_l96<<<(unsigned)((n * d + 255) / 256), 256, 0, s>>>(P.p[0], n, d, q, scale0, 0.0, 0, means + D,
                                                                    nullptr, nullptr, lin);
  else
#endif
    k_linearize_compact<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ivp_id, P, n, d, q, scale0, means + D, lin,
                                                                    loop_state);
  POF_CK(cudaGetLastError());
  LeafArgs a;
  if (int rc = make_args(n, d, q, qL_host, nullptr, nullptr, wl, flags, a)) return rc;
  a.Jc = lin;
  a.s0 = scale0;
  a.s1 = scale1;
  a.stop = loop_state;
  if (int rc = run_pass(s, ctx, flags, ll, a, wl, ws, N, x0_mean, x0_chol, means, chols, nullptr, nullptr, calibrate,
                        scalars))
    return rc;
  if (loop_state)
    k_crit<<<1, 1, 0, s>>>(scalars, loop_state, maxiters, cond ? *cond : cudaGraphConditionalHandle(), cond ? 1 : 0);
  return (int)cudaGetLastError();
}

int POF_SUFFIX(pof_ieks_iteration)(pof_stream_t s_, pof_ctx_t* ctx, uint32_t flags, int ivp_id,
                                   const double* params_host, int nparams, int64_t N, int d, int q, int64_t chunk_len,
                                   const double* qL_host, double scale0, double scale1, const real* x0_mean,
                                   const real* x0_chol, real* means, real* chols, int calibrate, real* scalars,
                                   void* ws_, size_t ws_bytes) {
  return ieks_iteration_impl((cudaStream_t)s_, ctx, flags, ivp_id, params_host, nparams, N, d, q, chunk_len, qL_host,
                             scale0, scale1, x0_mean, x0_chol, means, chols, calibrate, scalars, nullptr, 0.0, ws_,
                             ws_bytes);
}
// One step of the IEKS loop WITH the stopping rule on the device (the reference's lax.while_loop body + cond,
// solver.py:36-57): the host may enqueue several of these back to back without synchronising; once the rule has
// fired, the remaining calls are no-ops and `means`, `chols`, `scalars` stay those of the final iteration.
int POF_SUFFIX(pof_ieks_loop_step)(pof_stream_t s_, pof_ctx_t* ctx, uint32_t flags, int ivp_id,
                                   const double* params_host, int nparams, int64_t N, int d, int q, int64_t chunk_len,
                                   const double* qL_host, double scale0, double scale1, const real* x0_mean,
                                   const real* x0_chol, real* means, real* chols, int calibrate, real* scalars,
                                   real* loop_state, int64_t maxiters, void* ws_, size_t ws_bytes) {
  if (!loop_state) return POF_E_ARG;
  return ieks_iteration_impl((cudaStream_t)s_, ctx, flags, ivp_id, params_host, nparams, N, d, q, chunk_len, qL_host,
                             scale0, scale1, x0_mean, x0_chol, means, chols, calibrate, scalars, loop_state,
                             (double)maxiters, ws_, ws_bytes);
}

// The loop as ONE graph launch: body = one loop step, recorded by stream capture into the body graph of a WHILE
// conditional node; k_crit sets the node's condition.  (CUDA graphs in place of the reference's traced while_loop.)
int POF_SUFFIX(pof_ieks_loop_create)(pof_loop_t** out, pof_ctx_t* ctx, uint32_t flags, int ivp_id,
                                     const double* params_host, int nparams, int64_t N, int d, int q,
                                     int64_t chunk_len, const double* qL_host, double scale0, double scale1,
                                     const real* x0_mean, const real* x0_chol, real* means, real* chols, int calibrate,
                                     real* scalars, real* loop_state, int64_t maxiters, void* ws_, size_t ws_bytes) {
  if (!out || !loop_state) return POF_E_ARG;
  *out = nullptr;
  pof_loop* lp = new pof_loop();
  cudaStream_t cs = nullptr;  // capture needs a non-default stream of its own
  cudaGraphConditionalHandle cond;
  cudaGraphNodeParams np = {};
  cudaGraphNode_t node;
  cudaGraph_t body = nullptr, captured = nullptr;
  int rc = 0;
  cudaError_t e = cudaGraphCreate(&lp->graph, 0);
  if (e == cudaSuccess) e = cudaGraphConditionalHandleCreate(&cond, lp->graph, 1, cudaGraphCondAssignDefault);
  if (e == cudaSuccess) {
    np.type = cudaGraphNodeTypeConditional;
    np.conditional.handle = cond;
    np.conditional.type = cudaGraphCondTypeWhile;
    np.conditional.size = 1;
    e = cudaGraphAddNode(&node, lp->graph, nullptr, 0, &np);
  }
  if (e == cudaSuccess) body = np.conditional.phGraph_out[0];
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamBeginCaptureToGraph(cs, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
  if (e == cudaSuccess) {
    rc = ieks_iteration_impl(cs, ctx, flags, ivp_id, params_host, nparams, N, d, q, chunk_len, qL_host, scale0, scale1,
                             x0_mean, x0_chol, means, chols, calibrate, scalars, loop_state, (double)maxiters, ws_,
                             ws_bytes, &cond);
    e = cudaStreamEndCapture(cs, &captured);
  }
  if (e == cudaSuccess && rc == 0) e = cudaGraphInstantiate(&lp->exec, lp->graph, 0);
  if (cs) cudaStreamDestroy(cs);
  if (e != cudaSuccess || rc != 0) {
    (void)cudaGetLastError();
    if (lp->exec) cudaGraphExecDestroy(lp->exec);
    if (lp->graph) cudaGraphDestroy(lp->graph);
    delete lp;
    return rc != 0 ? rc : (int)e;
  }
  *out = lp;
  return 0;
}

#ifndef POF_F32
int pof_ieks_loop_launch(pof_loop_t* lp, pof_stream_t s) {
  if (!lp || !lp->exec) return POF_E_ARG;
  return (int)cudaGraphLaunch(lp->exec, (cudaStream_t)s);
}
void pof_ieks_loop_destroy(pof_loop_t* lp) {
  if (!lp) return;
  if (lp->exec) cudaGraphExecDestroy(lp->exec);
  if (lp->graph) cudaGraphDestroy(lp->graph);
  delete lp;
}
#endif

#ifndef POF_F32
int POF_SUFFIX(pof_sequential_eks)(pof_stream_t s_, uint32_t flags, int ivp_id, const double* params_host, int nparams,
                           int64_t N, int d, int q, const double* qL_host, double scale0, double scale1,
                           const real* x0_mean, const real* x0_chol, real* means, real* chols, real* scalars,
                           void* ws_, size_t ws_bytes) {
  cudaStream_t s = (cudaStream_t)s_;
  if (N < 2) return POF_E_ARG;
  IvpParams P;
  if (int rc = fill_params(ivp_id, params_host, nparams, d, P)) return rc;
  // one thread (d <= 4 templates) or, for larger states / POF_F_FAMILY_TILE, one CTA of the tile family
  const LeafLaunch* ll = (flags & POF_F_FAMILY_TILE) ? nullptr : seq_launch(d, q);
  if (!ll && tile_supported(d, q)) ll = tile_leaf_launch();
  if (!ll || !ll->seq_eks) return POF_E_UNSUPPORTED_DQ;
  WsLayout wl;
  wl.build(N - 1, d, q, N - 1);
  if (ws_bytes < wl.total * sizeof(real)) return POF_E_WORKSPACE;
  real* ws = (real*)ws_;
  LeafArgs a;
  if (int rc = make_args(N - 1, d, q, qL_host, nullptr, nullptr, wl, flags, a)) return rc;
  a.s0 = scale0;
  a.s1 = scale1;
  real* x0 = ws + wl.o_misc;
  k_pack_state<<<1, 128, 0, s>>>(wl.D, x0_mean, x0_chol, x0);
  POF_CK(ll->seq_eks(s, a, ivp_id, P.p, x0, ws + wl.o_kern, means, chols, ws + wl.o_sums));
  // scalars: NLL slot holds the reference's `ell` = +sum loglik (sequential path sign, filter.py:91)
  k_finalize_seq<<<1, 1, 0, s>>>(ws + wl.o_sums, (real)(N - 1), (real)d, scalars);
  return (int)cudaGetLastError();
}

#endif  // !POF_F32

// ---- time-sharded stages
static int shard_setup(int64_t n_loc, int d, int q, int64_t chunk_len, const double* qL_host, const real* H,
                       const real* c, const real* Jc, double s0, double s1, uint32_t flags, size_t ws_bytes,
                       const LeafLaunch*& ll, WsLayout& wl, LeafArgs& a) {
  if (n_loc < 1) return POF_E_ARG;
  ll = leaf_launch(d, q, flags);
  if (!ll) return POF_E_UNSUPPORTED_DQ;
  wl.build(n_loc, d, q, chunk_len);
  if (ws_bytes < wl.total * sizeof(real)) return POF_E_WORKSPACE;
  if (int rc = make_args(n_loc, d, q, qL_host, H, c, wl, flags, a)) return rc;
  if (Jc) {
    a.Jc = Jc;
    a.s0 = s0;
    a.s1 = s1;
  }
  return 0;
}
static int shard_a(cudaStream_t s, pof_ctx* ctx, uint32_t flags, int64_t n_loc, int d, int q, int64_t chunk_len,
                   const double* qL_host, const real* H, const real* c, const real* Jc, double s0, double s1,
                   real* carry_f, void* ws_, size_t ws_bytes) {
  const LeafLaunch* ll;
  WsLayout wl;
  LeafArgs a;
  if (int rc = shard_setup(n_loc, d, q, chunk_len, qL_host, H, c, Jc, s0, s1, flags, ws_bytes, ll, wl, a)) return rc;
  real* ws = (real*)ws_;
  if (int rc = zero_flags(s, wl, ws)) return rc;
  if (int rc = stage_a(s, ctx, flags, ll, a, wl, ws, true)) return rc;
  // the shard's carry = aggregate of all its chunks: the tree's root element, or with the hybrid sweep the final
  // element of the last Kogge-Stone node
  const bool hybrid = tree_for(ll, wl.D, flags) && !(flags & POF_F_TREE_PER_LEVEL) && use_hybrid(wl, flags);
  const real* total = hybrid ? ks_total(wl, ws) : ws + wl.o_fagg + wl.tl.off[wl.tl.nlev - 1] * wl.FE;
  POF_CK(cudaMemcpyAsync(carry_f, total, wl.FE * sizeof(real), cudaMemcpyDeviceToDevice, s));
  return 0;
}
static int shard_b(cudaStream_t s, pof_ctx* ctx, uint32_t flags, int64_t n_loc, int d, int q, int64_t chunk_len,
                   const double* qL_host, const real* H, const real* c, const real* Jc, double s0, double s1,
                   const real* state_in, real* fmeans, real* fchols, real* carry_s, real* state_end,
                   real* partials, void* ws_, size_t ws_bytes) {
  const LeafLaunch* ll;
  WsLayout wl;
  LeafArgs a;
  if (int rc = shard_setup(n_loc, d, q, chunk_len, qL_host, H, c, Jc, s0, s1, flags, ws_bytes, ll, wl, a)) return rc;
  real* ws = (real*)ws_;
  if (int rc = zero_flags(s, wl, ws)) return rc;
  if (int rc = stage_b(s, ctx, flags, ll, a, wl, ws, state_in, state_in + wl.D, fmeans, fchols, true, 1.0, 0, nullptr))
    return rc;
  POF_CK(cudaMemcpyAsync(carry_s, ws + wl.o_sagg + wl.tl.off[wl.tl.nlev - 1] * wl.SE, wl.SE * sizeof(real),
                         cudaMemcpyDeviceToDevice, s));
  POF_CK(cudaMemcpyAsync(state_end, ws + wl.o_send + (wl.CS - 1) * wl.ST, wl.ST * sizeof(real),
                         cudaMemcpyDeviceToDevice, s));
  POF_CK(cudaMemcpyAsync(partials, ws + wl.o_sums, 3 * sizeof(real), cudaMemcpyDeviceToDevice, s));
  return 0;
}

int POF_SUFFIX(pof_shard_stage_a)(pof_stream_t s_, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                          int64_t chunk_len, const double* qL_host, const real* H, const real* c, real* carry_f,
                          void* ws_, size_t ws_bytes) {
  return shard_a((cudaStream_t)s_, ctx, flags, n_loc, d, q, chunk_len, qL_host, H, c, nullptr, 0.0, 0.0, carry_f, ws_,
                 ws_bytes);
}
int POF_SUFFIX(pof_shard_stage_a_compact)(pof_stream_t s_, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                                  int64_t chunk_len, const double* qL_host, const real* Jc, double scale0,
                                  double scale1, real* carry_f, void* ws_, size_t ws_bytes) {
  return shard_a((cudaStream_t)s_, ctx, flags, n_loc, d, q, chunk_len, qL_host, nullptr, nullptr, Jc, scale0, scale1,
                 carry_f, ws_, ws_bytes);
}
int POF_SUFFIX(pof_shard_stage_b)(pof_stream_t s_, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                          int64_t chunk_len, const double* qL_host, const real* H, const real* c,
                          const real* state_in, real* fmeans, real* fchols, real* carry_s, real* state_end,
                          real* partials, void* ws_, size_t ws_bytes) {
  return shard_b((cudaStream_t)s_, ctx, flags, n_loc, d, q, chunk_len, qL_host, H, c, nullptr, 0.0, 0.0, state_in,
                 fmeans, fchols, carry_s, state_end, partials, ws_, ws_bytes);
}
int POF_SUFFIX(pof_shard_stage_b_compact)(pof_stream_t s_, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                                  int64_t chunk_len, const double* qL_host, const real* Jc, double scale0,
                                  double scale1, const real* state_in, real* fmeans, real* fchols,
                                  real* carry_s, real* state_end, real* partials, void* ws_, size_t ws_bytes) {
  return shard_b((cudaStream_t)s_, ctx, flags, n_loc, d, q, chunk_len, qL_host, nullptr, nullptr, Jc, scale0, scale1,
                 state_in, fmeans, fchols, carry_s, state_end, partials, ws_, ws_bytes);
}

int POF_SUFFIX(pof_shard_stage_c)(pof_stream_t s_, pof_ctx_t* ctx, uint32_t flags, int64_t n_loc, int d, int q,
                          int64_t chunk_len, const double* qL_host, const real* seed, int is_last_rank, int has_row0,
                          const real* cscale, real* means, real* chols, real* partials2, void* ws_,
                          size_t ws_bytes) {
  cudaStream_t s = (cudaStream_t)s_;
  (void)is_last_rank;
  const LeafLaunch* ll;
  WsLayout wl;
  LeafArgs a;
  if (int rc = shard_setup(n_loc, d, q, chunk_len, qL_host, nullptr, nullptr, nullptr, 0.0, 0.0, flags, ws_bytes, ll,
                           wl, a))
    return rc;
  real* ws = (real*)ws_;
  if (int rc = zero_flags(s, wl, ws)) return rc;
  // local row of state t' is t' - (1 - has_row0): shift the base pointers so that the kernels can index by t'
  const long shift = has_row0 ? 0 : 1;
  real* mb = means - shift * wl.D;
  real* cb = chols ? chols - shift * (long)wl.D * wl.D : nullptr;
  if (int rc = stage_c(s, ctx, flags, ll, a, wl, ws, seed, seed + wl.D, has_row0, cscale, mb, cb, nullptr, true))
    return rc;
  POF_CK(cudaMemcpyAsync(partials2, ws + wl.o_sums + 8, 2 * sizeof(real), cudaMemcpyDeviceToDevice, s));
  return 0;
}

// ---- fused rank-carry exchanges (register-resident family; else the caller uses the chain entry points below)
#ifndef POF_F32
int pof_shard_exchange_supported(int D, uint32_t flags) {
  return (!(flags & POF_F_FAMILY_TILE) && tree_launch(D) != nullptr) ? 1 : 0;
}
#endif  // !POF_F32

int POF_SUFFIX(pof_shard_exchange_filter)(pof_stream_t s, uint32_t flags, int D, int rank, int world, const real* gathered,
                                  int64_t stride, const real* x0_mean, const real* x0_chol, real* state_in,
                                  real* scratch) {
  if ((flags & POF_F_FAMILY_TILE) || tree_launch(D) == nullptr) return POF_E_UNSUPPORTED_DQ;
  if (rank < 0 || rank >= world) return POF_E_ARG;
  ExchangeArgs A = {};
  A.rank = rank;
  A.world = world;
  A.gathered = gathered;
  A.stride = stride;
  A.x0_mean = x0_mean;
  A.x0_chol = x0_chol;
  A.state_out = state_in;
  A.scratch = scratch;
  return (int)tree_launch(D)->fexchange((cudaStream_t)s, A);
}
int POF_SUFFIX(pof_shard_exchange_smooth)(pof_stream_t s, uint32_t flags, int D, int d, int rank, int world,
                                  int64_t n_steps_total, int calibrate, const real* gathered, int64_t stride,
                                  real* seed, real* scratch, real* cscale, real* scalars) {
  if ((flags & POF_F_FAMILY_TILE) || tree_launch(D) == nullptr) return POF_E_UNSUPPORTED_DQ;
  if (rank < 0 || rank >= world || !cscale) return POF_E_ARG;
  ExchangeArgs A = {};
  A.rank = rank;
  A.world = world;
  A.gathered = gathered;
  A.stride = stride;
  A.state_out = seed;
  A.scratch = scratch;
  A.n_obs = (real)n_steps_total;
  A.d_obs = (real)d;
  A.calibrate = calibrate;
  A.cscale = cscale;
  A.scalars = scalars;
  return (int)tree_launch(D)->sexchange((cudaStream_t)s, A);
}
// scalars[POF_S_OBJ], scalars[POF_S_NOT_CLOSE] <- sums over the ranks' (obj, not-close) pairs, in rank order
static __global__ void k_exchange_sums2(int world, const real* __restrict__ gathered, real* __restrict__ scalars) {
  real a = 0.0, b = 0.0;
  for (int r = 0; r < world; ++r) {
    a += gathered[2 * r];
    b += gathered[2 * r + 1];
  }
  scalars[POF_S_OBJ] = a;
  scalars[POF_S_NOT_CLOSE] = b;
}
int POF_SUFFIX(pof_shard_exchange_scalars)(pof_stream_t s, int world, const real* gathered, real* scalars) {
  k_exchange_sums2<<<1, 1, 0, (cudaStream_t)s>>>(world, gathered, scalars);
  return (int)cudaGetLastError();
}

#ifndef POF_F32
// ---- peer-memory exchanges (CUDA IPC): every rank owns one exchange area; peers store their payloads into it over
// NVLink and release a flag, the exchange kernel (k_exchange with A.p2p) pushes, waits and folds in ONE launch.
struct pof_p2p {
  int rank = 0, world = 1, D = 0;
  unsigned long long* area = nullptr;     // local, cudaMalloc'ed (IPC handles refer to whole allocations)
  unsigned long long* peer[8] = {nullptr};
  size_t words = 0;
  long stride[3] = {0, 0, 0};
  long slot_word[3] = {0, 0, 0};
};
enum { P2P_FLAG0 = 8, P2P_SLOTS = 32 };

int pof_p2p_create(int rank, int world, int D, pof_p2p_t** out, unsigned char* handle_out) {
  if (!out || !handle_out || world < 1 || world > 8 || rank < 0 || rank >= world || D < 1) return POF_E_ARG;
  pof_p2p* p = new pof_p2p();
  p->rank = rank;
  p->world = world;
  p->D = D;
  const long FE = 3L * D * D + 2 * D, SE = 2L * D * D + D, ST = (long)D * D + D;
  p->stride[0] = FE;
  p->stride[1] = SE + ST + 4;
  p->stride[2] = 2;
  long w = P2P_SLOTS;
  for (int x = 0; x < 3; ++x) {
    p->slot_word[x] = w;
    w += 2L * world * p->stride[x];  // two parities
    w = (w + 1) & ~1L;               // 16-byte aligned regions
  }
  p->words = (size_t)w;
  cudaError_t e = cudaMalloc(&p->area, p->words * 8);
  if (e == cudaSuccess) e = cudaMemset(p->area, 0, p->words * 8);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p->area);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    if (p->area) cudaFree(p->area);
    delete p;
    return (int)e;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == POF_P2P_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  p->peer[rank] = p->area;
  *out = p;
  return 0;
}
// handles: world x POF_P2P_HANDLE_BYTES in rank order (exchanged by the caller with any host-side mechanism)
int pof_p2p_connect(pof_p2p_t* p, const unsigned char* handles) {
  if (!p || !handles) return POF_E_ARG;
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * POF_P2P_HANDLE_BYTES, sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      return (int)e;
    }
    p->peer[r] = (unsigned long long*)ptr;
  }
  return 0;
}
void pof_p2p_destroy(pof_p2p_t* p) {
  if (!p) return;
  for (int r = 0; r < p->world; ++r)
    if (r != p->rank && p->peer[r]) cudaIpcCloseMemHandle(p->peer[r]);
  if (p->area) cudaFree(p->area);
  delete p;
}
// 0 = ok, 1 = an exchange timed out waiting for a peer (synchronises the device)
int pof_p2p_status(pof_p2p_t* p, int* status_host) {
  if (!p || !status_host) return POF_E_ARG;
  unsigned long long v = 0;
  cudaError_t e = cudaMemcpy(&v, p->area, 8, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return (int)e;
  *status_host = (int)v;
  return 0;
}
static void p2p_fill(ExchangeArgs& A, const pof_p2p* p, int x, const real* payload) {
  A.p2p = 1;
  A.rank = p->rank;
  A.world = p->world;
  A.payload = payload;
  A.stride = p->stride[x];
  for (int r = 0; r < 8; ++r) A.peer[r] = r < p->world ? p->peer[r] : nullptr;
  A.epoch_word = 1 + x;
  A.flag_word = P2P_FLAG0 + 8 * x;
  A.slot_word = p->slot_word[x];
}
int pof_p2p_exchange_filter_f64(pof_stream_t s, uint32_t flags, pof_p2p_t* p, const real* carry_f, const real* x0_mean,
                                const real* x0_chol, real* state_in, real* scratch) {
  if (!p) return POF_E_ARG;
  if ((flags & POF_F_FAMILY_TILE) || tree_launch(p->D) == nullptr) return POF_E_UNSUPPORTED_DQ;
  ExchangeArgs A = {};
  p2p_fill(A, p, 0, carry_f);
  A.x0_mean = x0_mean;
  A.x0_chol = x0_chol;
  A.state_out = state_in;
  A.scratch = scratch;
  return (int)tree_launch(p->D)->fexchange((cudaStream_t)s, A);
}
int pof_p2p_exchange_smooth_f64(pof_stream_t s, uint32_t flags, pof_p2p_t* p, int d, int64_t n_steps_total,
                                int calibrate, const real* payload, real* seed, real* scratch, real* cscale,
                                real* scalars) {
  if (!p || !cscale) return POF_E_ARG;
  if ((flags & POF_F_FAMILY_TILE) || tree_launch(p->D) == nullptr) return POF_E_UNSUPPORTED_DQ;
  ExchangeArgs A = {};
  p2p_fill(A, p, 1, payload);
  A.state_out = seed;
  A.scratch = scratch;
  A.n_obs = (real)n_steps_total;
  A.d_obs = (real)d;
  A.calibrate = calibrate;
  A.cscale = cscale;
  A.scalars = scalars;
  return (int)tree_launch(p->D)->sexchange((cudaStream_t)s, A);
}
// the third exchange: (obj, not-close) pairs of all ranks -> scalars[OBJ], scalars[NOT_CLOSE], summed in rank order
static __global__ void __launch_bounds__(32) k_exchange_scalars_p2p(ExchangeArgs A) {
  const real* g = p2p_exchange(A, 0, A.world);
  if ((threadIdx.x & 31) == 0) {
    real a = 0.0, b = 0.0;
    for (int r = 0; r < A.world; ++r) {
      a += __ldcg(g + 2 * r);
      b += __ldcg(g + 2 * r + 1);
    }
    A.scalars[POF_S_OBJ] = a;
    A.scalars[POF_S_NOT_CLOSE] = b;
  }
}
int pof_p2p_exchange_scalars_f64(pof_stream_t s, pof_p2p_t* p, const real* pair, real* scalars) {
  if (!p || !scalars) return POF_E_ARG;
  ExchangeArgs A = {};
  p2p_fill(A, p, 2, pair);
  A.scalars = scalars;
  k_exchange_scalars_p2p<<<1, 32, 0, (cudaStream_t)s>>>(A);
  return (int)cudaGetLastError();
}

int POF_SUFFIX(pof_filter_apply_chain)(pof_stream_t s, uint32_t flags, int D, int count, const real* state_in,
                               const real* elems, real* state_out, real* scratch) {
  if ((flags & POF_F_FAMILY_TILE) || tree_launch(D) == nullptr) {
    if (!tile_tree_supported(D)) return POF_E_UNSUPPORTED_DQ;
    return (int)tile_fchain((cudaStream_t)s, D, count, state_in, elems, state_out, scratch);
  }
  const int smem = coop_ws_doubles(D) * (int)sizeof(real);
  POF_CK(ensure_smem(k_filter_chain, smem));
  k_filter_chain<<<1, 32, smem, (cudaStream_t)s>>>(D, count, state_in, elems, state_out, scratch);
  return (int)cudaGetLastError();
}
int POF_SUFFIX(pof_smooth_apply_chain)(pof_stream_t s, uint32_t flags, int D, int count, const real* state_in,
                               const real* elems, real* state_out, real* scratch) {
  if ((flags & POF_F_FAMILY_TILE) || tree_launch(D) == nullptr) {
    if (!tile_tree_supported(D)) return POF_E_UNSUPPORTED_DQ;
    return (int)tile_schain((cudaStream_t)s, D, count, state_in, elems, state_out, scratch);
  }
  const int smem = coop_ws_doubles(D) * (int)sizeof(real);
  POF_CK(ensure_smem(k_smooth_chain, smem));
  k_smooth_chain<<<1, 32, smem, (cudaStream_t)s>>>(D, count, state_in, elems, state_out, scratch);
  return (int)cudaGetLastError();
}

#endif  // !POF_F32

int POF_SUFFIX(pof_prior_init)(pof_stream_t s, int64_t N, int d, int q, const double* qL_host, const real* ts,
                       const real* m0, real* means, real* chols) {
  if (N < 1 || d < 1 || q < 1 || q > 5) return POF_E_ARG;
  QLParam ql;
  for (int i = 0; i < 36; ++i) ql.v[i] = (i < (q + 1) * (q + 1)) ? qL_host[i] : 0.0;
  const long total = (long)N * d * (q + 1);
  k_prior_init<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)s>>>(N, d, q, ts, m0, ql, means, chols);
  return (int)cudaGetLastError();
}

int POF_SUFFIX(pof_project)(pof_stream_t s, int64_t N, int d, int q, double scale0, const real* mult_dev,
                    const real* means, const real* chols, real* ymean, real* ychol) {
  const long total = (long)N * d * (d * (q + 1) + 1);
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  k_project<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(N, d, q, scale0, mult_dev, means, chols, ymean, ychol);
  return (int)cudaGetLastError();
}

}  // extern "C"
