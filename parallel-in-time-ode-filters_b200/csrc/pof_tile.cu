// Kernels of the large-state ("tile") family: one CTA per time-chunk (leaf recursions) or per tree node (associative
// operators), runtime (d, q); the device code is pof_tile.cuh, shared verbatim with the host simulator.
#include <cuda_runtime.h>

#include "pof_launch.cuh"
#include "pof_tile.cuh"

namespace pof {

constexpr int TILE_THREADS = 256;
constexpr int TILE_SMEM_MAX = 227 * 1024;

static __device__ __forceinline__ TileLin make_lin(const LeafArgs& a) {
  TileLin lin;
  lin.H = a.H;
  lin.c = a.c;
  lin.Jc = a.Jc;
  lin.R = a.R;
  lin.s0 = a.s0;
  lin.s1 = a.s1;
  lin.F = a.F;
  lin.QL = a.QLd;
  lin.reg_sweeps = a.tile_reg;
  return lin;
}

__global__ void __launch_bounds__(TILE_THREADS, 1)
    k_tile_fold(LeafArgs a, double* __restrict__ fagg, double* __restrict__ faggm) {
  extern __shared__ __align__(16) double sm[];
  const long ch = blockIdx.x;
  const int D = a.d * (a.q + 1);
  const long FE = 3L * D * D + 2 * D;
  const long k0 = ch * a.L;
  const long k1 = (k0 + a.L < a.n) ? k0 + a.L : a.n;
  Team t;
  tile_fold(t, a.d, a.q, a.ql.v, make_lin(a), k0, k1, fagg + ch * FE, faggm ? faggm + ch * FE : nullptr, sm);
}

__global__ void __launch_bounds__(TILE_THREADS, 1)
    k_tile_scan(LeafArgs a, const double* __restrict__ fin, double* __restrict__ kern, double* __restrict__ send,
                double* __restrict__ part, double* __restrict__ fmeans, double* __restrict__ fchols) {
  extern __shared__ __align__(16) double sm[];
  const long ch = blockIdx.x;
  const int D = a.d * (a.q + 1);
  const long ST = (long)D * D + D;
  const long k0 = ch * a.L;
  const long k1 = (k0 + a.L < a.n) ? k0 + a.L : a.n;
  Team t;
  tile_scan(t, a.d, a.q, a.ql.v, make_lin(a), k0, k1, fin + ch * ST, kern, send + ch * ST, part + ch * 3, fmeans,
            fchols, sm);
}

__global__ void __launch_bounds__(TILE_THREADS, 1)
    k_tile_smooth(LeafArgs a, const double* __restrict__ sin, const double* __restrict__ kern, int emit_t0,
                  const double* __restrict__ cscale, double* __restrict__ means, double* __restrict__ chols,
                  double* __restrict__ part2) {
  extern __shared__ __align__(16) double sm[];
  const long ch = blockIdx.x;
  const int D = a.d * (a.q + 1);
  const long ST = (long)D * D + D;
  const long k0 = ch * a.L;
  const long k1 = (k0 + a.L < a.n) ? k0 + a.L : a.n;
  const double cs = cscale ? *cscale : 1.0;
  Team t;
  tile_smooth(t, a.d, a.q, a.ql.v, make_lin(a), k0, k1, ch == a.CS - 1, emit_t0 != 0, sin + ch * ST, kern, cs, means, chols,
              part2 + ch * 2, sm);
}

// sequential EKS: ONE CTA walks the whole grid (inherently sequential baseline path of the reference)
__global__ void __launch_bounds__(TILE_THREADS, 1)
    k_tile_seq_eks(LeafArgs a, TileEks eks, const double* __restrict__ x0, double* __restrict__ kern,
                   double* __restrict__ state_end, double* __restrict__ means, double* __restrict__ chols,
                   double* __restrict__ sums) {
  extern __shared__ __align__(16) double sm[];
  Team t;
  tile_seq_eks(t, a.d, a.q, a.ql.v, make_lin(a), eks, a.n, x0, kern, state_end, means, chols, sums, sm);
}

// ------------------------------------------------------------------------------------------------ tree sweeps
// same node conventions as the warp kernels in pof_api.cu (k_filter_up, ...), one CTA per parent node
__global__ void __launch_bounds__(TILE_THREADS, 1)
    k_tile_fup(int D, const double* __restrict__ child, long nchild, double* __restrict__ parent) {
  extern __shared__ __align__(16) double sm[];
  const long i = blockIdx.x;
  const long FE = 3L * D * D + 2 * D;
  const double* lc = child + 2 * i * FE;
  double* out = parent + i * FE;
  Team t;
  if (2 * i + 1 < nchild)
    tile_filter_combine(t, D, lc, lc + FE, out, sm, false);
  else
    t.each((int)FE, [&](int j) { out[j] = lc[j]; });
}
__global__ void __launch_bounds__(TILE_THREADS, 1)
    k_tile_fdown(int D, const double* __restrict__ pin, const double* __restrict__ cagg, long nchild,
                 double* __restrict__ cin) {
  extern __shared__ __align__(16) double sm[];
  const long i = blockIdx.x;
  const long FE = 3L * D * D + 2 * D, ST = (long)D * D + D;
  const double* p = pin + i * ST;
  double* c0 = cin + 2 * i * ST;
  Team t;
  t.each((int)ST, [&](int j) { c0[j] = p[j]; });
  if (2 * i + 1 < nchild) tile_filter_combine(t, D, p, cagg + 2 * i * FE, cin + (2 * i + 1) * ST, sm, true);
}
__global__ void __launch_bounds__(TILE_THREADS, 1)
    k_tile_sup(int D, const double* __restrict__ child, long nchild, double* __restrict__ parent) {
  extern __shared__ __align__(16) double sm[];
  const long i = blockIdx.x;
  const long SE = 2L * D * D + D;
  const double* lc = child + 2 * i * SE;
  double* out = parent + i * SE;
  Team t;
  if (2 * i + 1 < nchild)
    tile_smooth_combine(t, D, lc + SE, lc, out, sm, false);
  else
    t.each((int)SE, [&](int j) { out[j] = lc[j]; });
}
__global__ void __launch_bounds__(TILE_THREADS, 1)
    k_tile_sdown(int D, const double* __restrict__ pin, const double* __restrict__ cagg, long nchild,
                 double* __restrict__ cin) {
  extern __shared__ __align__(16) double sm[];
  const long i = blockIdx.x;
  const long SE = 2L * D * D + D, ST = (long)D * D + D;
  const double* p = pin + i * ST;
  Team t;
  if (2 * i + 1 < nchild) {
    double* c1 = cin + (2 * i + 1) * ST;
    t.each((int)ST, [&](int j) { c1[j] = p[j]; });
    tile_smooth_combine(t, D, p, cagg + (2 * i + 1) * SE, cin + 2 * i * ST, sm, true);
  } else {
    double* c0 = cin + 2 * i * ST;
    t.each((int)ST, [&](int j) { c0[j] = p[j]; });
  }
}
__global__ void __launch_bounds__(TILE_THREADS, 1)
    k_tile_chunkk(int D, const double* __restrict__ fin, const double* __restrict__ faggm, double* __restrict__ sagg) {
  extern __shared__ __align__(16) double sm[];
  const long i = blockIdx.x;
  const long FE = 3L * D * D + 2 * D, SE = 2L * D * D + D, ST = (long)D * D + D;
  Team t;
  tile_chunk_kernel(t, D, fin + i * ST, faggm + i * FE, sagg + i * SE, sm);
}
__global__ void __launch_bounds__(TILE_THREADS, 1)
    k_tile_fcomb(int D, const double* __restrict__ e1, const double* __restrict__ e2, double* __restrict__ out) {
  extern __shared__ __align__(16) double sm[];
  const long i = blockIdx.x;
  const long FE = 3L * D * D + 2 * D;
  Team t;
  tile_filter_combine(t, D, e1 + i * FE, e2 + i * FE, out + i * FE, sm, false);
}
__global__ void __launch_bounds__(TILE_THREADS, 1)
    k_tile_scomb(int D, const double* __restrict__ e1, const double* __restrict__ e2, double* __restrict__ out) {
  extern __shared__ __align__(16) double sm[];
  const long i = blockIdx.x;
  const long SE = 2L * D * D + D;
  Team t;
  tile_smooth_combine(t, D, e1 + i * SE, e2 + i * SE, out + i * SE, sm, false);
}
// sequential chains over a handful of rank carries (one CTA); ping-pong so that the last write lands in state_out
__global__ void __launch_bounds__(TILE_THREADS, 1)
    k_tile_fchain(int D, int count, const double* __restrict__ state_in, const double* __restrict__ elems,
                  double* __restrict__ state_out, double* __restrict__ scratch) {
  extern __shared__ __align__(16) double sm[];
  const long FE = 3L * D * D + 2 * D, ST = (long)D * D + D;
  Team t;
  const double* cur = state_in;
  for (int i = 0; i < count; ++i) {
    double* dst = ((count - 1 - i) % 2 == 0) ? state_out : scratch;
    tile_filter_combine(t, D, cur, elems + (long)i * FE, dst, sm, true);
    cur = dst;
  }
  if (count == 0) t.each((int)ST, [&](int j) { state_out[j] = state_in[j]; });
}
__global__ void __launch_bounds__(TILE_THREADS, 1)
    k_tile_schain(int D, int count, const double* __restrict__ state_in, const double* __restrict__ elems,
                  double* __restrict__ state_out, double* __restrict__ scratch) {
  extern __shared__ __align__(16) double sm[];
  const long SE = 2L * D * D + D, ST = (long)D * D + D;
  Team t;
  const double* cur = state_in;
  for (int i = 0; i < count; ++i) {
    double* dst = ((count - 1 - i) % 2 == 0) ? state_out : scratch;
    tile_smooth_combine(t, D, cur, elems + (long)(count - 1 - i) * SE, dst, sm, true);
    cur = dst;
  }
  if (count == 0) t.each((int)ST, [&](int j) { state_out[j] = state_in[j]; });
}

// ------------------------------------------------------------------------------------------------ launchers
static inline int bytes(int doubles) { return doubles * (int)sizeof(double); }

bool tile_supported(int d, int q) {
  if (d < 1 || q < 1 || q > 5) return false;
  const int D = d * (q + 1);
  const int need = bytes(tile_fold_smem_doubles(D, d));
  return need <= TILE_SMEM_MAX && bytes(tile_scan_smem_doubles(D, d)) <= TILE_SMEM_MAX &&
         bytes(tile_tree_smem_doubles(D)) <= TILE_SMEM_MAX;
}
bool tile_tree_supported(int D) { return D >= 1 && bytes(tile_tree_smem_doubles(D)) <= TILE_SMEM_MAX; }
int tile_ctas_per_sm(int d, int q) {
  const int D = d * (q + 1);
  int m = tile_fold_smem_doubles(D, d);
  if (tile_scan_smem_doubles(D, d) > m) m = tile_scan_smem_doubles(D, d);
  int c = TILE_SMEM_MAX / (bytes(m) + 1024);
  if (c < 1) c = 1;
  if (c > 2048 / TILE_THREADS) c = 2048 / TILE_THREADS;
  return c;
}

static cudaError_t tl_fold(cudaStream_t s, const LeafArgs& a, double* fagg, double* faggm) {
  const int D = a.d * (a.q + 1);
  const int sb = bytes(tile_fold_smem_doubles(D, a.d));
  if (cudaError_t e = ensure_smem(k_tile_fold, sb)) return e;
  k_tile_fold<<<(unsigned)a.CS, TILE_THREADS, sb, s>>>(a, fagg, faggm);
  return cudaGetLastError();
}
// the chunk smoothing elements come from the chunk-level op (tile_chunk_kernel): sagg must be null
static cudaError_t tl_scan(cudaStream_t s, const LeafArgs& a, const double* fin, double* kern, double* sagg,
                           double* send, double* part, double* fmeans, double* fchols) {
  if (sagg) return cudaErrorInvalidValue;
  const int D = a.d * (a.q + 1);
  const int sb = bytes(tile_scan_smem_doubles(D, a.d));
  if (cudaError_t e = ensure_smem(k_tile_scan, sb)) return e;
  k_tile_scan<<<(unsigned)a.CS, TILE_THREADS, sb, s>>>(a, fin, kern, send, part, fmeans, fchols);
  return cudaGetLastError();
}
static cudaError_t tl_smooth(cudaStream_t s, const LeafArgs& a, const double* sin, const double* kern, int emit_t0,
                             const double* cscale, double* means, double* chols, double* part2) {
  const int D = a.d * (a.q + 1);
  const int sb = bytes(tile_smooth_smem_doubles(D, a.d));
  if (cudaError_t e = ensure_smem(k_tile_smooth, sb)) return e;
  k_tile_smooth<<<(unsigned)a.CS, TILE_THREADS, sb, s>>>(a, sin, kern, emit_t0, cscale, means, chols, part2);
  return cudaGetLastError();
}
// kern: (n, NE) scratch; part: >= 5 doubles [sum -loglik, ssq_ref, ssq_proper, obj, -]; the D + D^2 doubles behind
// the packed x0 are used as scratch for the filtered end state
static cudaError_t tl_seq_eks(cudaStream_t s, const LeafArgs& a, int ivp_id, const double* params8, const double* x0,
                              double* kern, double* means, double* chols, double* part) {
  const int D = a.d * (a.q + 1);
  int dbl = tile_scan_smem_doubles(D, a.d);
  if (tile_smooth_smem_doubles(D, a.d) > dbl) dbl = tile_smooth_smem_doubles(D, a.d);
  const int sb = bytes(dbl);
  if (cudaError_t e = ensure_smem(k_tile_seq_eks, sb)) return e;
  TileEks eks;
  eks.ivp_id = ivp_id;
  for (int i = 0; i < 8; ++i) eks.P.p[i] = params8[i];
  double* state_end = const_cast<double*>(x0) + (D + D * D);
  k_tile_seq_eks<<<1, TILE_THREADS, sb, s>>>(a, eks, x0, kern, state_end, means, chols, part);
  return cudaGetLastError();
}
const LeafLaunch* tile_leaf_launch() {
  static const LeafLaunch l = {&tl_fold, &tl_scan, &tl_smooth, &tl_seq_eks, 0, 1, 1};
  return &l;
}

#define POF_TILE_TREE_LAUNCH(kernel, grid, ...)                        \
  do {                                                                 \
    if ((grid) <= 0) return cudaSuccess;                               \
    const int sb = bytes(tile_tree_smem_doubles(D));                   \
    if (cudaError_t e = ensure_smem(kernel, sb)) return e;             \
    kernel<<<(unsigned)(grid), TILE_THREADS, sb, s>>>(__VA_ARGS__);    \
    return cudaGetLastError();                                         \
  } while (0)

cudaError_t tile_fup(cudaStream_t s, int D, const double* child, long nchild, double* parent, long nparent) {
  POF_TILE_TREE_LAUNCH(k_tile_fup, nparent, D, child, nchild, parent);
}
cudaError_t tile_fdown(cudaStream_t s, int D, const double* pin, long nparent, const double* cagg, long nchild,
                       double* cin) {
  POF_TILE_TREE_LAUNCH(k_tile_fdown, nparent, D, pin, cagg, nchild, cin);
}
cudaError_t tile_sup(cudaStream_t s, int D, const double* child, long nchild, double* parent, long nparent) {
  POF_TILE_TREE_LAUNCH(k_tile_sup, nparent, D, child, nchild, parent);
}
cudaError_t tile_sdown(cudaStream_t s, int D, const double* pin, long nparent, const double* cagg, long nchild,
                       double* cin) {
  POF_TILE_TREE_LAUNCH(k_tile_sdown, nparent, D, pin, cagg, nchild, cin);
}
cudaError_t tile_chunkk(cudaStream_t s, int D, const double* fin, const double* faggm, double* sagg, long CS) {
  POF_TILE_TREE_LAUNCH(k_tile_chunkk, CS, D, fin, faggm, sagg);
}
cudaError_t tile_fcomb(cudaStream_t s, int D, long n, const double* e1, const double* e2, double* out) {
  POF_TILE_TREE_LAUNCH(k_tile_fcomb, n, D, e1, e2, out);
}
cudaError_t tile_scomb(cudaStream_t s, int D, long n, const double* e1, const double* e2, double* out) {
  POF_TILE_TREE_LAUNCH(k_tile_scomb, n, D, e1, e2, out);
}
cudaError_t tile_fchain(cudaStream_t s, int D, int count, const double* state_in, const double* elems,
                        double* state_out, double* scratch) {
  POF_TILE_TREE_LAUNCH(k_tile_fchain, 1, D, count, state_in, elems, state_out, scratch);
}
cudaError_t tile_schain(cudaStream_t s, int D, int count, const double* state_in, const double* elems,
                        double* state_out, double* scratch) {
  POF_TILE_TREE_LAUNCH(k_tile_schain, 1, D, count, state_in, elems, state_out, scratch);
}

}  // namespace pof
