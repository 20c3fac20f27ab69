// C ABI (include/pof_b200.h) of the B200-native parallel-in-time IEKS pass, plus the kernels that are not
// templated on (d, q): the tree sweeps over chunk carries (one warp per associative combine, pof_coop.cuh),
// the deterministic scalar reductions, the fused vector-field/Jacobian linearisation for the built-in IVPs
// and the final calibration + projection.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/pof_b200.h"
#include "pof_coop.cuh"
#include "pof_ivp.cuh"
#include "pof_launch.cuh"
#include "pof_pipeline.cuh"

namespace pof {

// register-resident tree sweeps when 2D <= 32 (POF_B200_TREE_IMPL=generic forces the shared-memory kernels)
static const TreeLaunch* tree_launch(int D) {
  const char* e = getenv("POF_B200_TREE_IMPL");
  if (e && (e[0] == 'g' || e[0] == 't')) return nullptr;  // generic (warp, shared memory) / tile (CTA per node)
  const TreeLaunch* t = tree_launch_a(D);
  if (!t) t = tree_launch_b(D);
  if (!t) t = tree_launch_c(D);
  return t;
}

// POF_B200_TREE_SWEEP=1: whole tree sweeps as single cooperative kernels with a grid barrier per level.  Measured on
// B200 (N = 2^20, 14 levels): 0.51 ms per filter tree against 0.33 ms with one graph-replayed launch per level -- a
// cooperative-groups grid barrier costs more than a kernel boundary inside a CUDA graph here, so the default stays
// one launch per level.
static bool tree_fused() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("POF_B200_TREE_SWEEP");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// Side stream on which the smoother's up-sweep runs concurrently with the filter scan: its inputs (the chunk-level
// smoothing elements) only depend on the fold and the filter down-sweep, and its ~14 latency-bound levels fit into
// the registers / shared memory the scan leaves free on every SM.  One stream + two events per device, created on
// first use (eagerly, i.e. before any graph capture); fork/join through events, so the pass stays stream-ordered on
// the caller's stream and is capturable.  POF_B200_OVERLAP=0 disables it.
struct SideStream {
  cudaStream_t s2 = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool ok = false;
};
static SideStream* side_stream() {
  static SideStream tab[64];
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("POF_B200_OVERLAP");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  SideStream& t = tab[dev];
  if (!t.ok) {
    if (cudaStreamCreateWithFlags(&t.s2, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&t.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&t.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    t.ok = true;
  }
  return &t;
}

// thread-per-chunk reference kernels (pof_leaf.cuh)
static const LeafLaunch* thread_launch(int d, int q) {
  switch (d) {
    case 1: return leaf_launch_d1(q);
    case 2: return leaf_launch_d2(q);
    case 3: return leaf_launch_d3(q);
    case 4: return leaf_launch_d4(q);
    default: return nullptr;
  }
}
// lane-cooperative performance kernels (pof_lane.cuh); D <= 32
static const LeafLaunch* lane_launch(int d, int q) {
  if (d * (q + 1) > 32) return nullptr;
  switch (d) {
    case 1: return lane_launch_d1(q);
    case 2: return lane_launch_d2(q);
    case 3: return lane_launch_d3(q);
    case 4: return lane_launch_d4(q);
    default: return nullptr;
  }
}
// two rows per lane (pof_lane2.cuh): the default where instantiated (D <= 16)
static const LeafLaunch* lane2_launch(int d, int q) {
  switch (d) {
    case 1: return lane2_launch_d1(q);
    case 2: return lane2_launch_d2(q);
    case 3: return lane2_launch_d3(q);
    case 4: return lane2_launch_d4(q);
    default: return nullptr;
  }
}
// POF_B200_LEAF_IMPL = thread | lane1 | lane2 selects a kernel family (debugging / cross-checks); default: lane2
// where available (and the register-resident tree ops are not disabled), else lane1, else thread
const LeafLaunch* leaf_launch(int d, int q) {
  const char* e = getenv("POF_B200_LEAF_IMPL");
  const bool want_tile = e && e[0] == 't' && e[1] == 'i';
  if (want_tile) return tile_supported(d, q) ? tile_leaf_launch() : nullptr;
  const bool want_thread = e && e[0] == 't';
  const bool want_lane1 = e && e[0] == 'l' && e[1] == 'a' && e[2] == 'n' && e[3] == 'e' && e[4] == '1';
  if (!want_thread) {
    if (!want_lane1 && tree_launch(d * (q + 1)) != nullptr) {
      const LeafLaunch* l2 = lane2_launch(d, q);
      if (l2) return l2;
    }
    const LeafLaunch* l = lane_launch(d, q);
    if (l) return l;
  }
  if (const LeafLaunch* l = thread_launch(d, q)) return l;
  // anything the (d, q)-templated families do not cover: the large-state tile family (D limited by shared memory)
  return tile_supported(d, q) ? tile_leaf_launch() : nullptr;
}

constexpr int TREE_WARPS = 4;  // max warps (= element pairs) per CTA in the tree kernels; fewer when D is large

// ------------------------------------------------------------------------------------------------ tree sweeps
// up:   parent[i] = op(child[2i], child[2i+1])           (copy if the second child is missing)
__global__ void __launch_bounds__(TREE_WARPS * 32)
    k_filter_up(int D, const double* __restrict__ child, long nchild, double* __restrict__ parent, long nparent) {
  extern __shared__ double sm[];
  const long gw = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= nparent) return;
  Warp w;
  const int FE = filter_elem_size(D);
  const double* lc = child + 2 * gw * FE;
  double* out = parent + gw * FE;
  if (2 * gw + 1 < nchild)
    filter_combine(w, D, lc, lc + FE, out, sm + (threadIdx.x >> 5) * coop_ws_doubles(D), false);
  else
    coop_copy(w, out, lc, FE);
}
// down (exclusive, state form): cin[2i] = pin[i] ; cin[2i+1] = op(state pin[i], cagg[2i])
__global__ void __launch_bounds__(TREE_WARPS * 32)
    k_filter_down(int D, const double* __restrict__ pin, long nparent, const double* __restrict__ cagg, long nchild,
                  double* __restrict__ cin) {
  extern __shared__ double sm[];
  const long gw = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= nparent) return;
  Warp w;
  const int FE = filter_elem_size(D), ST = state_size(D);
  const double* p = pin + gw * ST;
  coop_copy(w, cin + 2 * gw * ST, p, ST);
  if (2 * gw + 1 < nchild)
    filter_combine(w, D, p, cagg + 2 * gw * FE, cin + (2 * gw + 1) * ST,
                   sm + (threadIdx.x >> 5) * coop_ws_doubles(D), true);
}
// smoother up: parent[i] = op(e1 = later = child[2i+1], e2 = earlier = child[2i])
__global__ void __launch_bounds__(TREE_WARPS * 32)
    k_smooth_up(int D, const double* __restrict__ child, long nchild, double* __restrict__ parent, long nparent) {
  extern __shared__ double sm[];
  const long gw = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= nparent) return;
  Warp w;
  const int SE = smooth_elem_size(D);
  const double* lc = child + 2 * gw * SE;
  double* out = parent + gw * SE;
  if (2 * gw + 1 < nchild)
    smooth_combine(w, D, lc + SE, lc, out, sm + (threadIdx.x >> 5) * coop_ws_doubles(D), false);
  else
    coop_copy(w, out, lc, SE);
}
// smoother down (exclusive suffix, state form): cin[2i+1] = pin[i] ; cin[2i] = op(state pin[i], cagg[2i+1])
__global__ void __launch_bounds__(TREE_WARPS * 32)
    k_smooth_down(int D, const double* __restrict__ pin, long nparent, const double* __restrict__ cagg, long nchild,
                  double* __restrict__ cin) {
  extern __shared__ double sm[];
  const long gw = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= nparent) return;
  Warp w;
  const int SE = smooth_elem_size(D), ST = state_size(D);
  const double* p = pin + gw * ST;
  if (2 * gw + 1 < nchild) {
    coop_copy(w, cin + (2 * gw + 1) * ST, p, ST);
    smooth_combine(w, D, p, cagg + (2 * gw + 1) * SE, cin + 2 * gw * ST,
                   sm + (threadIdx.x >> 5) * coop_ws_doubles(D), true);
  } else {
    coop_copy(w, cin + 2 * gw * ST, p, ST);
  }
}
// batched operators (C-ABI test hooks / S3 seam): out[i] = op(e1[i], e2[i])
__global__ void __launch_bounds__(TREE_WARPS * 32)
    k_filter_combine_batched(int D, long n, const double* __restrict__ e1, const double* __restrict__ e2,
                             double* __restrict__ out) {
  extern __shared__ double sm[];
  const long gw = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= n) return;
  Warp w;
  const int FE = filter_elem_size(D);
  filter_combine(w, D, e1 + gw * FE, e2 + gw * FE, out + gw * FE, sm + (threadIdx.x >> 5) * coop_ws_doubles(D), false);
}
__global__ void __launch_bounds__(TREE_WARPS * 32)
    k_smooth_combine_batched(int D, long n, const double* __restrict__ e1, const double* __restrict__ e2,
                             double* __restrict__ out) {
  extern __shared__ double sm[];
  const long gw = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (gw >= n) return;
  Warp w;
  const int SE = smooth_elem_size(D);
  smooth_combine(w, D, e1 + gw * SE, e2 + gw * SE, out + gw * SE, sm + (threadIdx.x >> 5) * coop_ws_doubles(D), false);
}
// sequential chains over a handful of rank carries (one warp)
__global__ void __launch_bounds__(32)
    k_filter_chain(int D, int count, const double* __restrict__ state_in, const double* __restrict__ elems,
                   double* __restrict__ state_out, double* __restrict__ scratch) {
  extern __shared__ double sm[];
  Warp w;
  const int FE = filter_elem_size(D), ST = state_size(D);
  // ping-pong between state_out and scratch so that the last write lands in state_out
  const double* cur = state_in;
  for (int i = 0; i < count; ++i) {
    double* dst = ((count - 1 - i) % 2 == 0) ? state_out : scratch;
    filter_combine(w, D, cur, elems + (long)i * FE, dst, sm, true);
    w.sync();
    cur = dst;
  }
  if (count == 0) coop_copy(w, state_out, state_in, ST);
}
__global__ void __launch_bounds__(32)
    k_smooth_chain(int D, int count, const double* __restrict__ state_in, const double* __restrict__ elems,
                   double* __restrict__ state_out, double* __restrict__ scratch) {
  extern __shared__ double sm[];
  Warp w;
  const int SE = smooth_elem_size(D), ST = state_size(D);
  const double* cur = state_in;
  for (int i = 0; i < count; ++i) {
    double* dst = ((count - 1 - i) % 2 == 0) ? state_out : scratch;
    smooth_combine(w, D, cur, elems + (long)(count - 1 - i) * SE, dst, sm, true);
    w.sync();
    cur = dst;
  }
  if (count == 0) coop_copy(w, state_out, state_in, ST);
}

// ------------------------------------------------------------------------------------------------ reductions
// out[j] = sum_i part[i*NC + j], fixed summation order (deterministic), single CTA of 1024 threads: every thread
// accumulates its strided share of all NC components with independent loads, then a shuffle tree per warp and one
// over the 32 warp sums (the first version took 22 us per launch: 3 serial passes of dependent loads by 256 threads)
template <int NC>
__global__ void __launch_bounds__(1024) k_reduce_parts_t(const double* __restrict__ part, long cnt,
                                                         double* __restrict__ out) {
  __shared__ double sh[32][NC];
  double s[NC];
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
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      double v = sh[lane][j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
      if (lane == 0) out[j] = v;
    }
  }
}
static inline void reduce_parts(cudaStream_t s, const double* part, long cnt, int ncomp, double* out) {
  if (ncomp == 2) k_reduce_parts_t<2><<<1, 1024, 0, s>>>(part, cnt, out);
  else if (ncomp == 3) k_reduce_parts_t<3><<<1, 1024, 0, s>>>(part, cnt, out);
  else if (ncomp == 4) k_reduce_parts_t<4><<<1, 1024, 0, s>>>(part, cnt, out);
  else k_reduce_parts_t<5><<<1, 1024, 0, s>>>(part, cnt, out);
}
// scalars from the filter partial sums [nll, s1, s2] over n*d observations
__global__ void k_finalize_filter(const double* __restrict__ sums, double n, double d, int calibrate,
                                  double* __restrict__ scal) {
  const double ssq = sums[1] / n / d;
  scal[POF_S_NLL] = sums[0];
  scal[POF_S_SSQ] = ssq;
  scal[POF_S_SSQ_PROPER] = sums[2] / n / d;
  scal[POF_S_CSCALE] = calibrate ? sqrt(ssq) : 1.0;
}
__global__ void k_finalize_smooth(const double* __restrict__ sums, double* __restrict__ scal) {
  scal[POF_S_OBJ] = sums[0];
  scal[POF_S_NOT_CLOSE] = sums[1];
}
__global__ void k_finalize_seq(const double* __restrict__ sums, double n, double d, double* __restrict__ scal) {
  scal[POF_S_NLL] = -sums[0];
  scal[POF_S_SSQ] = sums[1] / n / d;
  scal[POF_S_SSQ_PROPER] = sums[2] / n / d;
  scal[POF_S_OBJ] = sums[3];
  scal[POF_S_NOT_CLOSE] = 0.0;
  scal[POF_S_CSCALE] = 1.0;
}
__global__ void k_pack_state(int D, const double* __restrict__ m, const double* __restrict__ L,
                             double* __restrict__ st) {
  for (int i = threadIdx.x; i < D + D * D; i += blockDim.x) st[i] = (i < D) ? m[i] : L[i - D];
}

// ------------------------------------------------------------------------------------------------ linearise
// one thread per step k: H_k = E1 - J E0, c_k = J y - f(y) at y = E0 m_{k+1}
__global__ void __launch_bounds__(256)
    k_linearize(int ivp_id, IvpParams P, long n, int d, int q, double scale0, double scale1,
                const double* __restrict__ means_t1, double* __restrict__ H, double* __restrict__ c) {
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int Q1 = q + 1, D = d * Q1;
  double y[4], f[4], J[16];
  for (int b = 0; b < d; ++b) y[b] = scale0 * means_t1[k * D + b * Q1];
  ivp_eval(ivp_id, P, y, f, J);
  for (int a = 0; a < d; ++a) {
    double ca = -f[a];
    for (int b = 0; b < d; ++b) ca = fma(J[a * d + b], y[b], ca);
    c[k * d + a] = ca;
    double* Hr = H + (k * d + a) * D;
    for (int j = 0; j < D; ++j) Hr[j] = 0.0;
    for (int b = 0; b < d; ++b) Hr[b * Q1] = -J[a * d + b] * scale0;
    Hr[a * Q1 + 1] += scale1;
  }
}
// compact linearisation: per step [J_f (d x d) | c (d)] with c = J_f y - f(y); the leaf kernels rebuild H on load
__global__ void __launch_bounds__(256)
    k_linearize_compact(int ivp_id, IvpParams P, long n, int d, int q, double scale0,
                        const double* __restrict__ means_t1, double* __restrict__ Jc) {
  const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int Q1 = q + 1, D = d * Q1;
  double y[4], f[4], J[16];
  for (int b = 0; b < d; ++b) y[b] = scale0 * means_t1[k * D + b * Q1];
  ivp_eval(ivp_id, P, y, f, J);
  double* o = Jc + k * (d * d + d);
  for (int a = 0; a < d; ++a) {
    double ca = -f[a];
    for (int b = 0; b < d; ++b) {
      ca = fma(J[a * d + b], y[b], ca);
      o[a * d + b] = J[a * d + b];
    }
    o[d * d + a] = ca;
  }
}
// Lorenz-96 (POF_IVP_LORENZ96; the larger-state problem of BASELINE config 5, not in the reference's ivp.py):
//   f_a = (y_{a+1} - y_{a-2}) y_{a-1} - y_a + F (cyclic), 4 <= d.  One thread per (step, component): row a of the
// Jacobian has the three entries d f_a / d y_{a+1} = y_{a-1}, d f_a / d y_{a-2} = -y_{a-1}, d f_a / d y_{a-1} =
// y_{a+1} - y_{a-2} and -1 on the diagonal.  dense != 0: H (n,d,D), c (n,d); else compact [J_f | c] per step.
__global__ void __launch_bounds__(256)
    k_linearize_l96(double forcing, long n, int d, int q, double scale0, double scale1, int dense,
                    const double* __restrict__ means_t1, double* __restrict__ H, double* __restrict__ c,
                    double* __restrict__ Jc) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * d) return;
  l96_linearize_row(forcing, idx / d, (int)(idx % d), d, q, scale0, scale1, dense, means_t1, H, c, Jc);
}
// built-in problem ids and the ODE dimension each one accepts
static bool ivp_dim_ok(int ivp_id, int d) {
  static const int dims[] = {1, 2, 2, 2, 3, 3, 4, 4, 4};
  if (ivp_id == POF_IVP_LORENZ96) return d >= 4 && d <= 64;
  return ivp_id >= 0 && ivp_id <= POF_IVP_HENONHEILES && dims[ivp_id] == d;
}

// ys = E0 states, with the (second) calibration multiplier of pof/solver.py:66-69 read from device memory
__global__ void __launch_bounds__(256)
    k_project(long N, int d, int q, double scale0, const double* __restrict__ mult, const double* __restrict__ means,
              const double* __restrict__ chols, double* __restrict__ ymean, double* __restrict__ ychol) {
  const int Q1 = q + 1, D = d * Q1;
  const long total = N * d * (D + 1);
  const double mu = mult ? *mult : 1.0;
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
    k_prior_init(long N, int d, int q, const double* __restrict__ ts, const double* __restrict__ m0, QLParam ql,
                 double* __restrict__ means, double* __restrict__ chols) {
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
  const double t = fabs(ts[k]);
  double fact[6];
  fact[0] = 1.0;
  for (int p = 1; p <= q; ++p) fact[p] = fact[p - 1] * p;
  // sv_j = t^(q-j+1/2) / (q-j)!,  svi_j = t^-(q-j+1/2) (q-j)!
  double acc = 0.0;
  for (int j = i; j < Q1; ++j) {
    const double svi = pow(t, -((double)(q - j) + 0.5)) * fact[q - j];
    acc = fma(binom(q - i, j - i), svi * m0[blk * Q1 + j], acc);
  }
  const double sv = pow(t, (double)(q - i) + 0.5) / fact[q - i];
  means[k * D + r] = sv * acc;
  if (chols) {
    double* row = chols + (k * D + r) * D;
    for (int c = 0; c < D; ++c) row[c] = 0.0;
    for (int j = 0; j <= i; ++j) row[blk * Q1 + j] = -sv * ql.v[i * Q1 + j];
  }
}

// 16 independent FMA chains per thread: saturates the FP64 pipe
__global__ void __launch_bounds__(256) k_dfma_peak(int iters, double* __restrict__ sink) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double m = 1.0 - 1e-12, c = 1e-13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
  }
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) t += a[i];
  if (t == 123.456) sink[threadIdx.x] = t;
}

// ------------------------------------------------------------------------------------------------ workspace
struct WsLayout {
  TreeLevels tl;
  long CS, L;
  int D, FE, SE, ST, NE;
  size_t o_lin, o_faggm, o_fagg, o_fin, o_sagg, o_sin, o_kern, o_send, o_part, o_part2, o_sums, o_misc, total;  // in doubles
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
    o_kern = take((size_t)L * NE * CS);
    o_send = take((size_t)CS * ST);
    o_part = take((size_t)CS * 3);
    o_part2 = take((size_t)CS * 2);
    o_sums = take(16);
    o_misc = take((size_t)2 * ST + 64);
    total = o;
  }
};

// warps per CTA such that the per-warp shared-memory workspaces fit in 227 KB (0 if even one does not fit)
static inline int tree_warps(int D) {
  const long per = (long)coop_ws_doubles(D) * (long)sizeof(double);
  long w = (227L * 1024L) / per;
  return (int)(w > TREE_WARPS ? TREE_WARPS : w);
}
static inline int tree_smem_bytes(int D) { return tree_warps(D) * coop_ws_doubles(D) * (int)sizeof(double); }
// Which carry-level kernels serve a pass when no register-resident TreeLaunch exists for D: the warp-per-combine
// shared-memory kernels above, or the CTA-per-node tile kernels (pof_tile.cu).  The tile leaves need the chunk-level
// smoothing op, which only the register-resident and the tile trees have; POF_B200_TREE_IMPL=tile forces the tile tree.
static bool use_tile_tree(int D, const LeafLaunch* ll) {
  const char* e = getenv("POF_B200_TREE_IMPL");
  if (e && e[0] == 't') return tile_tree_supported(D);
  if (tree_launch(D) != nullptr) return false;
  return (ll && ll->is_tile) || tree_warps(D) < 1;
}

template <class K>
static cudaError_t set_smem(K kernel, int bytes) {
  return ensure_smem(kernel, bytes);
}

// ---- optional per-segment device timing (used by bench.py for the roofline numbers): CUDA events recorded on the
// launching stream around each segment of a pass.  Off by default; global, not thread-safe (a measurement aid).
enum { SEG_FOLD = 0, SEG_FUP, SEG_FDOWN, SEG_SCAN, SEG_SUP, SEG_SDOWN, SEG_SMOOTH, SEG_COUNT };
struct Prof {
  bool on = false;
  static constexpr int MAXP = 4096;
  cudaEvent_t ev[MAXP][2];
  int seg[MAXP];
  int created = 0, used = 0;
  double acc[SEG_COUNT] = {0};
  long cnt[SEG_COUNT] = {0};
} g_prof;
struct ProfScope {
  int idx = -1;
  cudaStream_t s;
  ProfScope(int seg, cudaStream_t st) : s(st) {
    if (!g_prof.on || g_prof.used >= Prof::MAXP) return;
    idx = g_prof.used++;
    if (idx >= g_prof.created) {
      cudaEventCreate(&g_prof.ev[idx][0]);
      cudaEventCreate(&g_prof.ev[idx][1]);
      g_prof.created = idx + 1;
    }
    g_prof.seg[idx] = seg;
    cudaEventRecord(g_prof.ev[idx][0], s);
  }
  ~ProfScope() {
    if (idx >= 0) cudaEventRecord(g_prof.ev[idx][1], s);
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

static int make_args(long n, int d, int q, const double* qL_host, const double* H, const double* c,
                     const WsLayout& wl, LeafArgs& a) {
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
  {
    const char* e = getenv("POF_B200_TILE_SWEEP");
    a.tile_reg = (e && e[0] == 'r') ? 1 : 0;
  }
  a.d = d;
  a.q = q;
  a.s0 = a.s1 = 0.0;
  for (int i = 0; i < 36; ++i) a.ql.v[i] = 0.0;
  for (int i = 0; i < (q + 1) * (q + 1); ++i) a.ql.v[i] = qL_host[i];
  return 0;
}

static void fill_sweep(SweepArgs& sa, const WsLayout& wl, double* agg, double* st, int up_levels, int do_down) {
  sa.nlev = wl.tl.nlev;
  sa.up_begin = 0;
  sa.up_end = up_levels < 0 ? 0 : up_levels;
  sa.down_begin = do_down ? wl.tl.nlev - 1 : 0;
  sa.down_end = 0;
  sa.block_sync = 0;
  for (int l = 0; l < SweepArgs::MAXL; ++l) {
    sa.off[l] = l < wl.tl.nlev ? wl.tl.off[l] : 0;
    sa.sz[l] = l < wl.tl.nlev ? wl.tl.sz[l] : 0;
  }
  sa.agg = agg;
  sa.st = st;
  sa.root_m = sa.root_L = nullptr;
  sa.faggm = sa.fin = nullptr;
}
// The apex of a tree: the levels whose nodes fit into ONE CTA of the sweep kernel (cap nodes per level).  They run in
// a single launch with __syncthreads between levels: warm instruction cache and no kernel boundary for the ~4 levels
// at the top of the up-sweep and of the down-sweep, where a launch costs a full cold-start combine (~10 us) for a
// handful of nodes.  POF_B200_TREE_APEX=0 disables it.
static bool apex_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("POF_B200_TREE_APEX");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}
// first up-sweep level l (building level l+1) that fits the apex: sz[l+1] <= cap; up_total = number of up levels
static int apex_up_begin(const WsLayout& wl, int up_total, int cap) {
  int l = 0;
  while (l < up_total && wl.tl.sz[l + 1] > cap) ++l;
  return l;
}
// the down-sweep runs l = nlev-1 .. 1 (level l-1 from level l); apex while sz[l] <= cap: returns the level at which the
// per-level launches take over (apex covers l = nlev-1 .. ret+1)
static int apex_down_end(const WsLayout& wl, int cap) {
  int l = wl.tl.nlev - 1;
  while (l >= 1 && wl.tl.sz[l] <= cap) --l;
  return l;
}

// stage A: fold + filter up-sweep.  The rank's element ends at the tree root.
// need_root: the tree's root element is only consumed by the time-sharded form (it is the shard's carry)
static int stage_a(cudaStream_t s, const LeafLaunch* ll, const LeafArgs& a, const WsLayout& wl, double* ws,
                   bool need_root) {
  double* fagg = ws + wl.o_fagg;
  const bool tile = use_tile_tree(wl.D, ll);
  const bool pre = ll->has_pre_update && (tree_launch(wl.D) != nullptr || tile);
  {
    ProfScope ps(SEG_FOLD, s);
    POF_CK(ll->fold(s, a, fagg, pre ? ws + wl.o_faggm : nullptr));
  }
  const TreeLaunch* tl = tree_launch(wl.D);
  if (tl && tree_fused()) {
    // single GPU: the up-sweep runs fused with the down-sweep in stage_b; sharded: up-sweep incl. the root here
    if (!need_root) return 0;
    ProfScope ps(SEG_FUP, s);
    SweepArgs sa;
    fill_sweep(sa, wl, fagg, ws + wl.o_fin, wl.tl.nlev - 1, 0);
    POF_CK(tl->fsweep(s, sa));
    return 0;
  }
  ProfScope ps(SEG_FUP, s);
  const int smem = tree_smem_bytes(wl.D);
  const int tw = tree_warps(wl.D);
  if (!tl && !tile) {
    if (tw < 1) return POF_E_UNSUPPORTED_DQ;
    POF_CK(set_smem(k_filter_up, smem));
  }
  const int up_total = wl.tl.nlev - 1 - (need_root ? 0 : 1);
  const int a_up = (tl && apex_enabled()) ? apex_up_begin(wl, up_total, tl->fcap) : up_total;
  if (tl && a_up < up_total && need_root) {  // sharded: the apex of the up-sweep here (single GPU: in stage_b)
    SweepArgs sa;
    fill_sweep(sa, wl, fagg, ws + wl.o_fin, 0, 0);
    sa.up_begin = a_up;
    sa.up_end = up_total;
    sa.block_sync = 1;
    for (int l = 0; l < a_up; ++l)
      POF_CK(tl->fup(s, fagg + wl.tl.off[l] * wl.FE, wl.tl.sz[l], nullptr, fagg + wl.tl.off[l + 1] * wl.FE,
                     wl.tl.sz[l + 1]));
    POF_CK(tl->fsweep(s, sa));
    return (int)cudaGetLastError();
  }
  for (int l = 0; l < (tl ? a_up : up_total); ++l) {
    const long np = wl.tl.sz[l + 1];
    if (tl)
      POF_CK(tl->fup(s, fagg + wl.tl.off[l] * wl.FE, wl.tl.sz[l], nullptr, fagg + wl.tl.off[l + 1] * wl.FE, np));
    else if (tile)
      POF_CK(tile_fup(s, wl.D, fagg + wl.tl.off[l] * wl.FE, wl.tl.sz[l], fagg + wl.tl.off[l + 1] * wl.FE, np));
    else
      k_filter_up<<<(unsigned)((np + tw - 1) / tw), tw * 32, smem, s>>>(
          wl.D, fagg + wl.tl.off[l] * wl.FE, wl.tl.sz[l], fagg + wl.tl.off[l + 1] * wl.FE, np);
  }
  return (int)cudaGetLastError();
}
// stage B: filter down-sweep from the root's incoming state (already stored at fin[root]), scan, smoother up-sweep
// fuse_sdown (single GPU): the smoother's down-sweep runs in the same cooperative kernel as its up-sweep, seeded with
// the filtered state of the last chunk; stage_c must then be called with skip_down
static int stage_b(cudaStream_t s, const LeafLaunch* ll, const LeafArgs& a, const WsLayout& wl, double* ws,
                   double* fmeans, double* fchols, bool need_root, bool fuse_sdown = false) {
  double* fagg = ws + wl.o_fagg;
  double* fin = ws + wl.o_fin;
  double* sagg = ws + wl.o_sagg;
  const int smem = tree_smem_bytes(wl.D);
  const int tw = tree_warps(wl.D);
  const TreeLaunch* tl = tree_launch(wl.D);
  const bool tile = use_tile_tree(wl.D, ll);
  if (!tl && !tile) {
    if (tw < 1) return POF_E_UNSUPPORTED_DQ;
    POF_CK(set_smem(k_filter_down, smem));
    POF_CK(set_smem(k_smooth_up, smem));
  }
  const bool fused = tl && tree_fused();
  if (fused) {
    ProfScope ps(need_root ? SEG_FDOWN : SEG_FUP, s);
    SweepArgs sa;
    fill_sweep(sa, wl, fagg, fin, need_root ? 0 : wl.tl.nlev - 2, 1);
    POF_CK(tl->fsweep(s, sa));
  } else {
  ProfScope ps(SEG_FDOWN, s);
  int l_start = wl.tl.nlev - 1;
  if (tl && apex_enabled()) {
    // apex: (single GPU) the top of the up-sweep left over by stage_a, then the top of the down-sweep, one CTA
    const int up_total = wl.tl.nlev - 1 - (need_root ? 0 : 1);
    const int a_up = need_root ? up_total : apex_up_begin(wl, up_total, tl->fcap);
    const int a_dn = apex_down_end(wl, tl->fcap);
    if (a_up < up_total || a_dn < wl.tl.nlev - 1) {
      SweepArgs sa;
      fill_sweep(sa, wl, fagg, fin, 0, 0);
      sa.up_begin = a_up;
      sa.up_end = up_total;
      sa.down_begin = wl.tl.nlev - 1;
      sa.down_end = a_dn;
      sa.block_sync = 1;
      POF_CK(tl->fsweep(s, sa));
      l_start = a_dn;
    }
  }
  for (int l = l_start; l >= 1; --l) {
    const long np = wl.tl.sz[l];
    if (tl)
      POF_CK(tl->fdown(s, fin + wl.tl.off[l] * wl.ST, wl.tl.sz[l - 1], fagg + wl.tl.off[l - 1] * wl.FE,
                       fin + wl.tl.off[l - 1] * wl.ST, np));
    else if (tile)
      POF_CK(tile_fdown(s, wl.D, fin + wl.tl.off[l] * wl.ST, np, fagg + wl.tl.off[l - 1] * wl.FE, wl.tl.sz[l - 1],
                        fin + wl.tl.off[l - 1] * wl.ST));
    else
      k_filter_down<<<(unsigned)((np + tw - 1) / tw), tw * 32, smem, s>>>(
          wl.D, fin + wl.tl.off[l] * wl.ST, np, fagg + wl.tl.off[l - 1] * wl.FE, wl.tl.sz[l - 1],
          fin + wl.tl.off[l - 1] * wl.ST);
  }
  }
  POF_CK(cudaGetLastError());
  const bool pre = ll->has_pre_update && (tl != nullptr || tile);
  SideStream* side = (pre && !fused && tl) ? side_stream() : nullptr;
  if (side) {
    // chunk-level smoothing elements, then fork: smoother up-sweep on the side stream || filter scan on s
    POF_CK(tl->chunkk(s, fin, wl.CS, ws + wl.o_faggm, sagg, wl.CS));
    POF_CK(cudaEventRecord(side->fork, s));
    POF_CK(cudaStreamWaitEvent(side->s2, side->fork, 0));
    {
      ProfScope ps(SEG_SUP, side->s2);
      const int up_total = wl.tl.nlev - 1 - (need_root ? 0 : 1);
      const int a_up = apex_enabled() ? apex_up_begin(wl, up_total, tl->scap) : up_total;
      for (int l = 0; l < a_up; ++l)
        POF_CK(tl->sup(side->s2, sagg + wl.tl.off[l] * wl.SE, wl.tl.sz[l], nullptr, sagg + wl.tl.off[l + 1] * wl.SE,
                       wl.tl.sz[l + 1]));
      if (a_up < up_total) {
        SweepArgs sa;
        fill_sweep(sa, wl, sagg, ws + wl.o_sin, 0, 0);
        sa.up_begin = a_up;
        sa.up_end = up_total;
        sa.block_sync = 1;
        POF_CK(tl->ssweep(side->s2, sa));
      }
    }
    POF_CK(cudaEventRecord(side->join, side->s2));
    {
      ProfScope ps(SEG_SCAN, s);
      POF_CK(ll->scan(s, a, fin, ws + wl.o_kern, nullptr, ws + wl.o_send, ws + wl.o_part, fmeans, fchols));
    }
    POF_CK(cudaStreamWaitEvent(s, side->join, 0));
    reduce_parts(s, ws + wl.o_part, wl.CS, 3, ws + wl.o_sums);
    return (int)cudaGetLastError();
  }
  {
    ProfScope ps(SEG_SCAN, s);
    POF_CK(ll->scan(s, a, fin, ws + wl.o_kern, pre ? nullptr : sagg, ws + wl.o_send, ws + wl.o_part, fmeans, fchols));
  }
  if (fused && pre) {
    ProfScope ps(SEG_SUP, s);
    SweepArgs sa;
    fill_sweep(sa, wl, sagg, ws + wl.o_sin, need_root ? wl.tl.nlev - 1 : wl.tl.nlev - 2, fuse_sdown ? 1 : 0);
    sa.faggm = ws + wl.o_faggm;
    sa.fin = fin;
    if (fuse_sdown) {
      sa.root_m = ws + wl.o_send + (wl.CS - 1) * wl.ST;
      sa.root_L = sa.root_m + wl.D;
    }
    POF_CK(tl->ssweep(s, sa));
    reduce_parts(s, ws + wl.o_part, wl.CS, 3, ws + wl.o_sums);
    return (int)cudaGetLastError();
  }
  // chunk-level smoothing elements straight from (incoming state, filtering element before its last update)
  if (pre) {
    if (tl)
      POF_CK(tl->chunkk(s, fin, wl.CS, ws + wl.o_faggm, sagg, wl.CS));
    else
      POF_CK(tile_chunkk(s, wl.D, fin, ws + wl.o_faggm, sagg, wl.CS));
  }
  ProfScope ps(SEG_SUP, s);
  for (int l = 0; l + 1 < wl.tl.nlev - (need_root ? 0 : 1); ++l) {
    const long np = wl.tl.sz[l + 1];
    if (tl)
      POF_CK(tl->sup(s, sagg + wl.tl.off[l] * wl.SE, wl.tl.sz[l], nullptr, sagg + wl.tl.off[l + 1] * wl.SE, np));
    else if (tile)
      POF_CK(tile_sup(s, wl.D, sagg + wl.tl.off[l] * wl.SE, wl.tl.sz[l], sagg + wl.tl.off[l + 1] * wl.SE, np));
    else
      k_smooth_up<<<(unsigned)((np + tw - 1) / tw), tw * 32, smem, s>>>(
          wl.D, sagg + wl.tl.off[l] * wl.SE, wl.tl.sz[l], sagg + wl.tl.off[l + 1] * wl.SE, np);
  }
  reduce_parts(s, ws + wl.o_part, wl.CS, 3, ws + wl.o_sums);
  return (int)cudaGetLastError();
}
// stage C: smoother down-sweep from the seed (already stored at sin[root]) + smoother scan
static int stage_c(cudaStream_t s, const LeafLaunch* ll, const LeafArgs& a, const WsLayout& wl, double* ws,
                   int emit_t0, const double* cscale, double* means, double* chols, bool skip_down = false) {
  double* sagg = ws + wl.o_sagg;
  double* sin_ = ws + wl.o_sin;
  const int smem = tree_smem_bytes(wl.D);
  const int tw = tree_warps(wl.D);
  const TreeLaunch* tl = tree_launch(wl.D);
  const bool tile = use_tile_tree(wl.D, ll);
  if (!tl && !tile) {
    if (tw < 1) return POF_E_UNSUPPORTED_DQ;
    POF_CK(set_smem(k_smooth_down, smem));
  }
  if (skip_down) {
  } else if (tl && tree_fused()) {
    ProfScope ps(SEG_SDOWN, s);
    SweepArgs sa;
    fill_sweep(sa, wl, sagg, sin_, 0, 1);
    POF_CK(tl->ssweep(s, sa));
  } else {
  ProfScope ps(SEG_SDOWN, s);
  int l_start = wl.tl.nlev - 1;
  if (tl && apex_enabled()) {
    const int a_dn = apex_down_end(wl, tl->scap);
    if (a_dn < wl.tl.nlev - 1) {
      SweepArgs sa;
      fill_sweep(sa, wl, sagg, sin_, 0, 0);
      sa.down_begin = wl.tl.nlev - 1;
      sa.down_end = a_dn;
      sa.block_sync = 1;
      POF_CK(tl->ssweep(s, sa));
      l_start = a_dn;
    }
  }
  for (int l = l_start; l >= 1; --l) {
    const long np = wl.tl.sz[l];
    if (tl)
      POF_CK(tl->sdown(s, sin_ + wl.tl.off[l] * wl.ST, wl.tl.sz[l - 1], sagg + wl.tl.off[l - 1] * wl.SE,
                       sin_ + wl.tl.off[l - 1] * wl.ST, np));
    else if (tile)
      POF_CK(tile_sdown(s, wl.D, sin_ + wl.tl.off[l] * wl.ST, np, sagg + wl.tl.off[l - 1] * wl.SE, wl.tl.sz[l - 1],
                        sin_ + wl.tl.off[l - 1] * wl.ST));
    else
      k_smooth_down<<<(unsigned)((np + tw - 1) / tw), tw * 32, smem, s>>>(
          wl.D, sin_ + wl.tl.off[l] * wl.ST, np, sagg + wl.tl.off[l - 1] * wl.SE, wl.tl.sz[l - 1],
          sin_ + wl.tl.off[l - 1] * wl.ST);
  }
  }
  POF_CK(cudaGetLastError());
  {
    ProfScope ps(SEG_SMOOTH, s);
    POF_CK(ll->smooth(s, a, sin_, ws + wl.o_kern, emit_t0, cscale, means, chols, ws + wl.o_part2));
  }
  reduce_parts(s, ws + wl.o_part2, wl.CS, 2, ws + wl.o_sums + 8);
  return (int)cudaGetLastError();
}

}  // namespace pof

using namespace pof;

extern "C" {

int pof_supported(int d, int q) { return leaf_launch(d, q) != nullptr ? 1 : 0; }

void pof_profile_enable(int on) {
  g_prof.on = on != 0;
  g_prof.used = 0;
  for (int i = 0; i < SEG_COUNT; ++i) {
    g_prof.acc[i] = 0.0;
    g_prof.cnt[i] = 0;
  }
}
// synchronises the device; ms_out[7] = accumulated milliseconds of [fold, filter_up, filter_down, scan, smooth_up,
// smooth_down, smooth], count_out[7] = number of timed segments of each kind
int pof_profile_read(double* ms_out, int64_t* count_out) {
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return (int)e;
  for (int i = 0; i < g_prof.used; ++i) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, g_prof.ev[i][0], g_prof.ev[i][1]) == cudaSuccess) {
      g_prof.acc[g_prof.seg[i]] += ms;
      g_prof.cnt[g_prof.seg[i]] += 1;
    }
  }
  g_prof.used = 0;
  for (int i = 0; i < SEG_COUNT; ++i) {
    ms_out[i] = g_prof.acc[i];
    count_out[i] = g_prof.cnt[i];
  }
  return 0;
}
// kernels launched by one pof_linear_filtsmooth_f64 call
int64_t pof_launches_per_pass(int64_t N, int d, int q, int64_t chunk_len) {
  WsLayout wl;
  wl.build(N - 1, d, q, chunk_len);
  const LeafLaunch* ll = leaf_launch(d, q);
  if (ll && ll->has_pre_update && tree_launch(wl.D) && tree_fused())
    return 3 /*leaf*/ + 2 /*cooperative tree sweeps*/ + 1 /*pack*/ + 2 /*reduce*/ + 2 /*finalize*/;
  const int up_total = wl.tl.nlev >= 2 ? wl.tl.nlev - 2 : 0;  // the root combine is skipped on one GPU
  const int down_total = wl.tl.nlev - 1;
  const TreeLaunch* tl = tree_launch(wl.D);
  int64_t tree = 2 * (int64_t)(up_total + down_total);
  if (tl && apex_enabled()) {  // the top levels of each sweep run in one single-CTA launch (stage_a / stage_b / stage_c)
    const int fu = apex_up_begin(wl, up_total, tl->fcap), fd = apex_down_end(wl, tl->fcap);
    const int su = apex_up_begin(wl, up_total, tl->scap), sd = apex_down_end(wl, tl->scap);
    // the smoother's up-sweep has its apex only on the side-stream path (two-rows-per-lane leaves, overlap enabled)
    const bool sup_apex = ll && ll->has_pre_update && side_stream() != nullptr;
    tree = fu + fd + ((fu < up_total || fd < down_total) ? 1 : 0)  // filter: per-level ups and downs, one apex
           + (sup_apex ? su + ((su < up_total) ? 1 : 0) : up_total) + sd + ((sd < down_total) ? 1 : 0);
  }
  const int64_t chunkk = (ll && ll->has_pre_update && (tl || use_tile_tree(wl.D, ll))) ? 1 : 0;
  return 3 /*leaf*/ + tree /*tree sweeps*/ + chunkk /*chunk smoothing elements*/ + 1 /*pack*/ + 2 /*reduce*/ +
         2 /*finalize*/;
}

// FP64 FMA throughput of this device (TFLOP/s), measured with a register-resident DFMA loop: the roofline
// denominator for the FP64-bound kernels (MEASURED_PEAKS.json holds no FP64 number)
int pof_measure_dfma_tflops(pof_stream_t s_, double* tflops_out) {
  cudaStream_t s = (cudaStream_t)s_;
  int dev = 0, sms = 0;
  POF_CK(cudaGetDevice(&dev));
  POF_CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  double* sink = nullptr;
  POF_CK(cudaMalloc(&sink, sizeof(double) * 1024));
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
  const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
  *tflops_out = flops / (best * 1e-3) / 1e12;
  return 0;
}

int64_t pof_default_chunk_len(int64_t N, int d, int q, int sm_count) {
  if (sm_count <= 0) sm_count = 148;
  const LeafLaunch* ll = leaf_launch(d, q);
  const int cpw = ll ? ll->chunks_per_warp : 32;
  const int64_t n = N - 1;
  // ~8 resident warps per SM, `cpw` chunks per warp; tile family: one chunk per resident CTA
  const int64_t target = (ll && ll->is_tile) ? (int64_t)sm_count * tile_ctas_per_sm(d, q) : (int64_t)sm_count * 8 * cpw;
  int64_t L = (n + target - 1) / target;
  if (L < 4) L = 4;
  return L;
}

int pof_supported_tile(int d, int q) { return tile_supported(d, q) ? 1 : 0; }
int64_t pof_default_chunk_len_tile(int64_t N, int d, int q, int sm_count) {
  if (sm_count <= 0) sm_count = 148;
  const int64_t n = N - 1;
  const int64_t target = (int64_t)sm_count * tile_ctas_per_sm(d, q);
  int64_t L = (n + target - 1) / target;
  return L < 4 ? 4 : L;
}

size_t pof_workspace_bytes(int64_t N, int d, int q, int64_t chunk_len) {
  WsLayout wl;
  wl.build(N - 1, d, q, chunk_len);
  return wl.total * sizeof(double);
}

int pof_filter_combine_f64(pof_stream_t s, int64_t n, int D, const double* e1, const double* e2, double* out) {
  if (n <= 0) return 0;
  if (const TreeLaunch* tl = tree_launch(D)) return (int)tl->fcomb((cudaStream_t)s, e1, n, e2, out, n);
  if (use_tile_tree(D, nullptr)) return (int)tile_fcomb((cudaStream_t)s, D, n, e1, e2, out);
  const int smem = tree_smem_bytes(D);
  const int tw = tree_warps(D);
  if (tw < 1) return POF_E_UNSUPPORTED_DQ;
  POF_CK(set_smem(k_filter_combine_batched, smem));
  k_filter_combine_batched<<<(unsigned)((n + tw - 1) / tw), tw * 32, smem, (cudaStream_t)s>>>(D, n, e1, e2, out);
  return (int)cudaGetLastError();
}
int pof_smooth_combine_f64(pof_stream_t s, int64_t n, int D, const double* e1, const double* e2, double* out) {
  if (n <= 0) return 0;
  if (const TreeLaunch* tl = tree_launch(D)) return (int)tl->scomb((cudaStream_t)s, e1, n, e2, out, n);
  if (use_tile_tree(D, nullptr)) return (int)tile_scomb((cudaStream_t)s, D, n, e1, e2, out);
  const int smem = tree_smem_bytes(D);
  const int tw = tree_warps(D);
  if (tw < 1) return POF_E_UNSUPPORTED_DQ;
  POF_CK(set_smem(k_smooth_combine_batched, smem));
  k_smooth_combine_batched<<<(unsigned)((n + tw - 1) / tw), tw * 32, smem, (cudaStream_t)s>>>(D, n, e1, e2, out);
  return (int)cudaGetLastError();
}

int pof_linearize_ivp_f64(pof_stream_t s, int ivp_id, const double* params_host, int nparams, int64_t n, int d, int q,
                          double scale0, double scale1, const double* means_t1, double* H, double* c) {
  if (ivp_id < 0 || ivp_id > POF_IVP_LORENZ96) return POF_E_IVP;
  if (nparams > 8 || !ivp_dim_ok(ivp_id, d)) return POF_E_ARG;
  IvpParams P;
  for (int i = 0; i < 8; ++i) P.p[i] = (i < nparams) ? params_host[i] : 0.0;
  if (n <= 0) return 0;
  if (ivp_id == POF_IVP_LORENZ96) {
    k_linearize_l96<<<(unsigned)((n * d + 255) / 256), 256, 0, (cudaStream_t)s>>>(P.p[0], n, d, q, scale0, scale1, 1,
                                                                                  means_t1, H, c, nullptr);
    return (int)cudaGetLastError();
  }
  k_linearize<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(ivp_id, P, n, d, q, scale0, scale1, means_t1,
                                                                        H, c);
  return (int)cudaGetLastError();
}

static int run_pass(cudaStream_t s, const LeafLaunch* ll, const LeafArgs& a, const WsLayout& wl, double* ws, int64_t N,
                    int d, const double* x0_mean, const double* x0_chol, double* means, double* chols,
                    double* fmeans, double* fchols, int calibrate, double* scalars);

int pof_linearize_ivp_compact_f64(pof_stream_t s, int ivp_id, const double* params_host, int nparams, int64_t n, int d,
                                  int q, double scale0, const double* means_t1, double* Jc) {
  if (ivp_id < 0 || ivp_id > POF_IVP_LORENZ96) return POF_E_IVP;
  if (nparams > 8 || !ivp_dim_ok(ivp_id, d)) return POF_E_ARG;
  IvpParams P;
  for (int i = 0; i < 8; ++i) P.p[i] = (i < nparams) ? params_host[i] : 0.0;
  if (n <= 0) return 0;
  if (ivp_id == POF_IVP_LORENZ96) {
    k_linearize_l96<<<(unsigned)((n * d + 255) / 256), 256, 0, (cudaStream_t)s>>>(P.p[0], n, d, q, scale0, 0.0, 0,
                                                                                  means_t1, nullptr, nullptr, Jc);
    return (int)cudaGetLastError();
  }
  k_linearize_compact<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)s>>>(ivp_id, P, n, d, q, scale0, means_t1,
                                                                                Jc);
  return (int)cudaGetLastError();
}

int pof_linear_filtsmooth_f64(pof_stream_t s_, int64_t N, int d, int q, int64_t chunk_len, const double* qL_host,
                              const double* x0_mean, const double* x0_chol, const double* H, const double* c,
                              double* means, double* chols, double* fmeans, double* fchols, int calibrate,
                              double* scalars, void* ws_, size_t ws_bytes) {
  cudaStream_t s = (cudaStream_t)s_;
  if (N < 2) return POF_E_ARG;
  const LeafLaunch* ll = leaf_launch(d, q);
  if (!ll) return POF_E_UNSUPPORTED_DQ;
  WsLayout wl;
  wl.build(N - 1, d, q, chunk_len);
  if (ws_bytes < wl.total * sizeof(double)) return POF_E_WORKSPACE;
  double* ws = (double*)ws_;
  LeafArgs a;
  int rc = make_args(N - 1, d, q, qL_host, H, c, wl, a);
  if (rc) return rc;
  return run_pass(s, ll, a, wl, ws, N, d, x0_mean, x0_chol, means, chols, fmeans, fchols, calibrate, scalars);
}

// general linear-Gaussian model: always the tile family (the only leaves that carry the D-column posterior factor and
// read dense per-step transition models)
int pof_linear_filtsmooth_general_f64(pof_stream_t s_, int64_t N, int d, int q, int64_t chunk_len,
                                      const double* qL_host, const double* F, const double* QL, const double* x0_mean,
                                      const double* x0_chol, const double* H, const double* c, const double* cholR,
                                      double* means, double* chols, double* fmeans, double* fchols, int calibrate,
                                      double* scalars, void* ws_, size_t ws_bytes) {
  cudaStream_t s = (cudaStream_t)s_;
  if (N < 2) return POF_E_ARG;
  if ((F == nullptr) != (QL == nullptr)) return POF_E_ARG;
  if (!F && !qL_host) return POF_E_ARG;
  if (!tile_supported(d, q)) return POF_E_UNSUPPORTED_DQ;
  const LeafLaunch* ll = tile_leaf_launch();
  WsLayout wl;
  wl.build(N - 1, d, q, chunk_len);
  if (ws_bytes < wl.total * sizeof(double)) return POF_E_WORKSPACE;
  double* ws = (double*)ws_;
  LeafArgs a;
  double ql_dummy[36] = {0.0};
  int rc = make_args(N - 1, d, q, qL_host ? qL_host : ql_dummy, H, c, wl, a);
  if (rc) return rc;
  a.R = cholR;
  a.F = F;
  a.QLd = QL;
  return run_pass(s, ll, a, wl, ws, N, d, x0_mean, x0_chol, means, chols, fmeans, fchols, calibrate, scalars);
}

int pof_ieks_iteration_f64(pof_stream_t s_, int ivp_id, const double* params_host, int nparams, int64_t N, int d,
                           int q, int64_t chunk_len, const double* qL_host, double scale0, double scale1,
                           const double* x0_mean, const double* x0_chol, double* means, double* chols, int calibrate,
                           double* scalars, void* ws_, size_t ws_bytes) {
  cudaStream_t s = (cudaStream_t)s_;
  if (N < 2) return POF_E_ARG;
  if (ivp_id < 0 || ivp_id > POF_IVP_LORENZ96) return POF_E_IVP;
  if (nparams > 8 || !ivp_dim_ok(ivp_id, d)) return POF_E_ARG;
  const LeafLaunch* ll = leaf_launch(d, q);
  if (!ll) return POF_E_UNSUPPORTED_DQ;
  WsLayout wl;
  wl.build(N - 1, d, q, chunk_len);
  if (ws_bytes < wl.total * sizeof(double)) return POF_E_WORKSPACE;
  double* ws = (double*)ws_;
  IvpParams P;
  for (int i = 0; i < 8; ++i) P.p[i] = (i < nparams) ? params_host[i] : 0.0;
  const long n = N - 1;
  const int D = wl.D;
  double* lin = ws + wl.o_lin;
  LeafArgs a;
  int rc;
  if (ivp_id == POF_IVP_LORENZ96) {
    if (!ll->has_pre_update) return POF_E_UNSUPPORTED_DQ;  // only the tile (and lane) leaves rebuild H from [J_f | c]
    k_linearize_l96<<<(unsigned)((n * d + 255) / 256), 256, 0, s>>>(P.p[0], n, d, q, scale0, 0.0, 0, means + D,
                                                                    nullptr, nullptr, lin);
    rc = make_args(n, d, q, qL_host, nullptr, nullptr, wl, a);
    a.Jc = lin;
    a.s0 = scale0;
    a.s1 = scale1;
  } else if (ll->has_pre_update) {  // lane kernels: compact linearisation, H rebuilt on load
    k_linearize_compact<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ivp_id, P, n, d, q, scale0, means + D, lin);
    rc = make_args(n, d, q, qL_host, nullptr, nullptr, wl, a);
    a.Jc = lin;
    a.s0 = scale0;
    a.s1 = scale1;
  } else {
    k_linearize<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(ivp_id, P, n, d, q, scale0, scale1, means + D, lin,
                                                            lin + (size_t)n * d * D);
    rc = make_args(n, d, q, qL_host, lin, lin + (size_t)n * d * D, wl, a);
  }
  if (rc) return rc;
  POF_CK(cudaGetLastError());
  return run_pass(s, ll, a, wl, ws, N, d, x0_mean, x0_chol, means, chols, nullptr, nullptr, calibrate, scalars);
}

}  // extern "C"

static int run_pass(cudaStream_t s, const LeafLaunch* ll, const LeafArgs& a, const WsLayout& wl, double* ws, int64_t N,
                    int d, const double* x0_mean, const double* x0_chol, double* means, double* chols,
                    double* fmeans, double* fchols, int calibrate, double* scalars) {
  int rc = stage_a(s, ll, a, wl, ws, false);
  if (rc) return rc;
  // root's incoming state = x0
  k_pack_state<<<1, 128, 0, s>>>(wl.D, x0_mean, x0_chol, ws + wl.o_fin + wl.tl.off[wl.tl.nlev - 1] * wl.ST);
  if (fmeans) {
    POF_CK(cudaMemcpyAsync(fmeans, x0_mean, wl.D * sizeof(double), cudaMemcpyDeviceToDevice, s));
    POF_CK(cudaMemcpyAsync(fchols, x0_chol, wl.D * wl.D * sizeof(double), cudaMemcpyDeviceToDevice, s));
  }
  const bool fuse_sdown = ll->has_pre_update && tree_launch(wl.D) != nullptr && tree_fused();
  rc = stage_b(s, ll, a, wl, ws, fmeans, fchols, false, fuse_sdown);
  if (rc) return rc;
  k_finalize_filter<<<1, 1, 0, s>>>(ws + wl.o_sums, (double)(N - 1), (double)d, calibrate, scalars);
  // terminal smoothing state = filtered state at the last time point
  if (!fuse_sdown)
    POF_CK(cudaMemcpyAsync(ws + wl.o_sin + wl.tl.off[wl.tl.nlev - 1] * wl.ST, ws + wl.o_send + (wl.CS - 1) * wl.ST,
                           wl.ST * sizeof(double), cudaMemcpyDeviceToDevice, s));
  rc = stage_c(s, ll, a, wl, ws, 1, scalars + POF_S_CSCALE, means, chols, fuse_sdown);
  if (rc) return rc;
  k_finalize_smooth<<<1, 1, 0, s>>>(ws + wl.o_sums + 8, scalars);
  return (int)cudaGetLastError();
}

extern "C" {

int pof_sequential_eks_f64(pof_stream_t s_, int ivp_id, const double* params_host, int nparams, int64_t N, int d,
                           int q, const double* qL_host, double scale0, double scale1, const double* x0_mean,
                           const double* x0_chol, double* means, double* chols, double* scalars, void* ws_,
                           size_t ws_bytes) {
  cudaStream_t s = (cudaStream_t)s_;
  if (N < 2) return POF_E_ARG;
  if (ivp_id < 0 || ivp_id > POF_IVP_LORENZ96) return POF_E_IVP;
  if (nparams > 8 || !ivp_dim_ok(ivp_id, d)) return POF_E_ARG;
  // one thread (d <= 4 templates) or, for larger states / POF_B200_LEAF_IMPL=tile, one CTA of the tile family
  const LeafLaunch* ll = leaf_launch(d, q);
  if (!ll || !ll->is_tile) ll = thread_launch(d, q);
  if (!ll || !ll->seq_eks) return POF_E_UNSUPPORTED_DQ;
  WsLayout wl;
  wl.build(N - 1, d, q, N - 1);
  if (ws_bytes < wl.total * sizeof(double)) return POF_E_WORKSPACE;
  double* ws = (double*)ws_;
  LeafArgs a;
  int rc = make_args(N - 1, d, q, qL_host, nullptr, nullptr, wl, a);
  if (rc) return rc;
  a.s0 = scale0;
  a.s1 = scale1;
  double p8[8];
  for (int i = 0; i < 8; ++i) p8[i] = (i < nparams) ? params_host[i] : 0.0;
  double* x0 = ws + wl.o_misc;
  k_pack_state<<<1, 128, 0, s>>>(wl.D, x0_mean, x0_chol, x0);
  POF_CK(ll->seq_eks(s, a, ivp_id, p8, x0, ws + wl.o_kern, means, chols, ws + wl.o_sums));
  // scalars: NLL slot holds the reference's `ell` = +sum loglik (sequential path sign, filter.py:91)
  k_finalize_seq<<<1, 1, 0, s>>>(ws + wl.o_sums, (double)(N - 1), (double)d, scalars);
  return (int)cudaGetLastError();
}

static int shard_a(cudaStream_t s, int64_t n_loc, int d, int q, int64_t chunk_len, const double* qL_host,
                   const double* H, const double* c, const double* Jc, double s0, double s1, double* carry_f,
                   void* ws_, size_t ws_bytes);
static int shard_b(cudaStream_t s, int64_t n_loc, int d, int q, int64_t chunk_len, const double* qL_host,
                   const double* H, const double* c, const double* Jc, double s0, double s1, const double* state_in,
                   double* fmeans, double* fchols, double* carry_s, double* state_end, double* partials, void* ws_,
                   size_t ws_bytes);

int pof_shard_stage_a_f64(pof_stream_t s_, int64_t n_loc, int d, int q, int64_t chunk_len, const double* qL_host,
                          const double* H, const double* c, double* carry_f, void* ws_, size_t ws_bytes) {
  return shard_a((cudaStream_t)s_, n_loc, d, q, chunk_len, qL_host, H, c, nullptr, 0.0, 0.0, carry_f, ws_, ws_bytes);
}
int pof_shard_stage_a_compact_f64(pof_stream_t s_, int64_t n_loc, int d, int q, int64_t chunk_len,
                                  const double* qL_host, const double* Jc, double scale0, double scale1,
                                  double* carry_f, void* ws_, size_t ws_bytes) {
  return shard_a((cudaStream_t)s_, n_loc, d, q, chunk_len, qL_host, nullptr, nullptr, Jc, scale0, scale1, carry_f, ws_,
                 ws_bytes);
}
int pof_shard_stage_b_f64(pof_stream_t s_, int64_t n_loc, int d, int q, int64_t chunk_len, const double* qL_host,
                          const double* H, const double* c, const double* state_in, double* fmeans, double* fchols,
                          double* carry_s, double* state_end, double* partials, void* ws_, size_t ws_bytes) {
  return shard_b((cudaStream_t)s_, n_loc, d, q, chunk_len, qL_host, H, c, nullptr, 0.0, 0.0, state_in, fmeans, fchols,
                 carry_s, state_end, partials, ws_, ws_bytes);
}
int pof_shard_stage_b_compact_f64(pof_stream_t s_, int64_t n_loc, int d, int q, int64_t chunk_len,
                                  const double* qL_host, const double* Jc, double scale0, double scale1,
                                  const double* state_in, double* fmeans, double* fchols, double* carry_s,
                                  double* state_end, double* partials, void* ws_, size_t ws_bytes) {
  return shard_b((cudaStream_t)s_, n_loc, d, q, chunk_len, qL_host, nullptr, nullptr, Jc, scale0, scale1, state_in,
                 fmeans, fchols, carry_s, state_end, partials, ws_, ws_bytes);
}

}  // extern "C"

static int shard_a(cudaStream_t s, int64_t n_loc, int d, int q, int64_t chunk_len, const double* qL_host,
                   const double* H, const double* c, const double* Jc, double s0, double s1, double* carry_f,
                   void* ws_, size_t ws_bytes) {
  if (n_loc < 1) return POF_E_ARG;
  const LeafLaunch* ll = leaf_launch(d, q);
  if (!ll) return POF_E_UNSUPPORTED_DQ;
  WsLayout wl;
  wl.build(n_loc, d, q, chunk_len);
  if (ws_bytes < wl.total * sizeof(double)) return POF_E_WORKSPACE;
  double* ws = (double*)ws_;
  LeafArgs a;
  int rc = make_args(n_loc, d, q, qL_host, H, c, wl, a);
  if (rc) return rc;
  if (Jc) {
    if (!ll->has_pre_update) return POF_E_ARG;  // only the lane kernels read the compact form
    a.Jc = Jc;
    a.s0 = s0;
    a.s1 = s1;
  }
  rc = stage_a(s, ll, a, wl, ws, true);
  if (rc) return rc;
  POF_CK(cudaMemcpyAsync(carry_f, ws + wl.o_fagg + wl.tl.off[wl.tl.nlev - 1] * wl.FE, wl.FE * sizeof(double),
                         cudaMemcpyDeviceToDevice, s));
  return 0;
}

static int shard_b(cudaStream_t s, int64_t n_loc, int d, int q, int64_t chunk_len, const double* qL_host,
                   const double* H, const double* c, const double* Jc, double s0, double s1, const double* state_in,
                   double* fmeans, double* fchols, double* carry_s, double* state_end, double* partials, void* ws_,
                   size_t ws_bytes) {
  const LeafLaunch* ll = leaf_launch(d, q);
  if (!ll) return POF_E_UNSUPPORTED_DQ;
  WsLayout wl;
  wl.build(n_loc, d, q, chunk_len);
  if (ws_bytes < wl.total * sizeof(double)) return POF_E_WORKSPACE;
  double* ws = (double*)ws_;
  LeafArgs a;
  int rc = make_args(n_loc, d, q, qL_host, H, c, wl, a);
  if (rc) return rc;
  if (Jc) {
    if (!ll->has_pre_update) return POF_E_ARG;
    a.Jc = Jc;
    a.s0 = s0;
    a.s1 = s1;
  }
  POF_CK(cudaMemcpyAsync(ws + wl.o_fin + wl.tl.off[wl.tl.nlev - 1] * wl.ST, state_in, wl.ST * sizeof(double),
                         cudaMemcpyDeviceToDevice, s));
  rc = stage_b(s, ll, a, wl, ws, fmeans, fchols, true);
  if (rc) return rc;
  POF_CK(cudaMemcpyAsync(carry_s, ws + wl.o_sagg + wl.tl.off[wl.tl.nlev - 1] * wl.SE, wl.SE * sizeof(double),
                         cudaMemcpyDeviceToDevice, s));
  POF_CK(cudaMemcpyAsync(state_end, ws + wl.o_send + (wl.CS - 1) * wl.ST, wl.ST * sizeof(double),
                         cudaMemcpyDeviceToDevice, s));
  POF_CK(cudaMemcpyAsync(partials, ws + wl.o_sums, 3 * sizeof(double), cudaMemcpyDeviceToDevice, s));
  return 0;
}

extern "C" {

int pof_shard_stage_c_f64(pof_stream_t s_, int64_t n_loc, int d, int q, int64_t chunk_len, const double* qL_host,
                          const double* seed, int is_last_rank, int has_row0, const double* cscale, double* means,
                          double* chols, double* partials2, void* ws_, size_t ws_bytes) {
  cudaStream_t s = (cudaStream_t)s_;
  (void)is_last_rank;
  const LeafLaunch* ll = leaf_launch(d, q);
  if (!ll) return POF_E_UNSUPPORTED_DQ;
  WsLayout wl;
  wl.build(n_loc, d, q, chunk_len);
  if (ws_bytes < wl.total * sizeof(double)) return POF_E_WORKSPACE;
  double* ws = (double*)ws_;
  LeafArgs a;
  int rc = make_args(n_loc, d, q, qL_host, nullptr, nullptr, wl, a);
  if (rc) return rc;
  POF_CK(cudaMemcpyAsync(ws + wl.o_sin + wl.tl.off[wl.tl.nlev - 1] * wl.ST, seed, wl.ST * sizeof(double),
                         cudaMemcpyDeviceToDevice, s));
  // local row of state t' is t' - (1 - has_row0): shift the base pointers so that the kernels can index by t'
  const long shift = has_row0 ? 0 : 1;
  double* mb = means - shift * wl.D;
  double* cb = chols ? chols - shift * (long)wl.D * wl.D : nullptr;
  rc = stage_c(s, ll, a, wl, ws, has_row0, cscale, mb, cb);
  if (rc) return rc;
  POF_CK(cudaMemcpyAsync(partials2, ws + wl.o_sums + 8, 2 * sizeof(double), cudaMemcpyDeviceToDevice, s));
  return 0;
}

int pof_filter_apply_chain_f64(pof_stream_t s, int D, int count, const double* state_in, const double* elems,
                               double* state_out, double* scratch) {
  if (use_tile_tree(D, nullptr))
    return (int)tile_fchain((cudaStream_t)s, D, count, state_in, elems, state_out, scratch);
  const int smem = coop_ws_doubles(D) * (int)sizeof(double);
  POF_CK(set_smem(k_filter_chain, smem));
  k_filter_chain<<<1, 32, smem, (cudaStream_t)s>>>(D, count, state_in, elems, state_out, scratch);
  return (int)cudaGetLastError();
}
int pof_smooth_apply_chain_f64(pof_stream_t s, int D, int count, const double* state_in, const double* elems,
                               double* state_out, double* scratch) {
  if (use_tile_tree(D, nullptr))
    return (int)tile_schain((cudaStream_t)s, D, count, state_in, elems, state_out, scratch);
  const int smem = coop_ws_doubles(D) * (int)sizeof(double);
  POF_CK(set_smem(k_smooth_chain, smem));
  k_smooth_chain<<<1, 32, smem, (cudaStream_t)s>>>(D, count, state_in, elems, state_out, scratch);
  return (int)cudaGetLastError();
}

int pof_prior_init_f64(pof_stream_t s, int64_t N, int d, int q, const double* qL_host, const double* ts,
                       const double* m0, double* means, double* chols) {
  if (N < 1 || d < 1 || q < 1 || q > 5) return POF_E_ARG;
  QLParam ql;
  for (int i = 0; i < 36; ++i) ql.v[i] = (i < (q + 1) * (q + 1)) ? qL_host[i] : 0.0;
  const long total = (long)N * d * (q + 1);
  k_prior_init<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)s>>>(N, d, q, ts, m0, ql, means, chols);
  return (int)cudaGetLastError();
}

int pof_project_f64(pof_stream_t s, int64_t N, int d, int q, double scale0, const double* mult_dev,
                    const double* means, const double* chols, double* ymean, double* ychol) {
  const long total = (long)N * d * (d * (q + 1) + 1);
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  k_project<<<(unsigned)blocks, 256, 0, (cudaStream_t)s>>>(N, d, q, scale0, mult_dev, means, chols, ymean, ychol);
  return (int)cudaGetLastError();
}

}  // extern "C"
