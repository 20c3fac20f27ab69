// Leaf kernels, performance path v2: G lanes per chunk, two matrix rows per lane (pof_lane2.cuh).
#pragma once
#include "pof_lane2.cuh"
#include "pof_launch.cuh"

namespace POF_NS {

constexpr int LANE2_WARPS = 4;

template <int d, int q>
struct Lane2Setup {
  using LN = Lane2<d, q>;
  static __device__ __forceinline__ long chunk_of_thread() {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return ((long)blockIdx.x * LANE2_WARPS + warp) * LN::GPW + lane / LN::G;
  }
  static __device__ __forceinline__ real* smem_of_thread(real* sm) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return sm + (warp * LN::GPW + lane / LN::G) * LN::SM_GROUP;
  }
  static constexpr int smem_bytes() { return LANE2_WARPS * LN::GPW * LN::SM_GROUP * (int)sizeof(real); }
  // smoother with bulk-copy staging: one staging block per chunk behind the groups' work areas (16-byte aligned).
  // Only used where TWO CTAs per SM still fit (the kernels run one wave of 2 CTAs x 4 warps per SM).
  static __host__ __device__ constexpr int stage_offset() { return ((LANE2_WARPS * LN::GPW * LN::SM_GROUP + 1) / 2) * 2; }
  static constexpr int smem_bytes_staged() {
    return (stage_offset() + LANE2_WARPS * LN::GPW * LN::STG) * (int)sizeof(real);
  }
  static constexpr bool staged_ok() { return LN::STG_OK && 2 * (smem_bytes_staged() + 1024) <= 227 * 1024; }
  static unsigned grid(long CS) {
    const long per_block = (long)LANE2_WARPS * LN::GPW;
    return (unsigned)((CS + per_block - 1) / per_block);
  }
};

template <int d, int q>
__global__ void __launch_bounds__(LANE2_WARPS * 32)
    k_lane2_fold(LeafArgs a, real* __restrict__ fagg, real* __restrict__ faggm) {
  extern __shared__ __align__(16) real sm[];
  using LN = Lane2<d, q>;
  if (a.stop && *a.stop != real(0)) return;  // the device-side IEKS loop has ended
  const long ch = Lane2Setup<d, q>::chunk_of_thread();
  if (ch >= a.CS) return;
  typename LN::Ctx cx;
  LN::init_ctx(cx, Lane2Setup<d, q>::smem_of_thread(sm), a.ql.v);
  const long k0 = ch * a.L;
  const long k1 = (k0 + a.L < a.n) ? k0 + a.L : a.n;
  constexpr int FE = 3 * LN::D * LN::D + 2 * LN::D;
  const typename LN::Lin lin = {a.H, a.c, a.Jc, a.s0, a.s1};
  LN::fold(cx, k0, k1, lin, fagg + ch * FE, faggm ? faggm + ch * FE : nullptr);
}

template <int d, int q>
__global__ void __launch_bounds__(LANE2_WARPS * 32)
    k_lane2_scan(LeafArgs a, const real* __restrict__ fin, real* __restrict__ kern, real* __restrict__ send,
                 real* __restrict__ part, real* __restrict__ fmeans, real* __restrict__ fchols) {
  extern __shared__ __align__(16) real sm[];
  using LN = Lane2<d, q>;
  if (a.stop && *a.stop != real(0)) return;
  const long ch = Lane2Setup<d, q>::chunk_of_thread();
  if (ch >= a.CS) return;
  typename LN::Ctx cx;
  LN::init_ctx(cx, Lane2Setup<d, q>::smem_of_thread(sm), a.ql.v);
  constexpr int D = LN::D, ST = D * D + D;
  const long k0 = ch * a.L;
  const long k1 = (k0 + a.L < a.n) ? k0 + a.L : a.n;
  const typename LN::Lin lin = {a.H, a.c, a.Jc, a.s0, a.s1};
  LN::scan(cx, k0, k1, lin, fin + ch * ST, kern, send + ch * ST, part + ch * 3, fmeans, fchols);
}

// STAGED: the next step's backward kernel is brought into shared memory by the bulk-copy (TMA) engine (pof_lane2.cuh)
template <int d, int q, bool STAGED>
__global__ void __launch_bounds__(LANE2_WARPS * 32)
    k_lane2_smooth(LeafArgs a, const real* __restrict__ sin, const real* __restrict__ kern, int emit_t0,
                   const real* __restrict__ cscale, real* __restrict__ means, real* __restrict__ chols,
                   real* __restrict__ part2) {
  extern __shared__ __align__(16) real sm[];
  using LN = Lane2<d, q>;
  if (a.stop && *a.stop != real(0)) return;  // (means, chols and the partial sums stay those of the final iteration)
  const long ch = Lane2Setup<d, q>::chunk_of_thread();
  if (ch >= a.CS) return;
  typename LN::Ctx cx;
  LN::init_ctx(cx, Lane2Setup<d, q>::smem_of_thread(sm), a.ql.v);
  if constexpr (STAGED) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    LN::stage_init(cx, sm + Lane2Setup<d, q>::stage_offset() + (warp * LN::GPW + lane / LN::G) * LN::STG);
  }
  constexpr int D = LN::D, ST = D * D + D;
  real qinv[LN::Q1];
#pragma unroll
  for (int i = 0; i < LN::Q1; ++i) qinv[i] = 1.0 / a.ql.v[i * LN::Q1 + i];
  const long k0 = ch * a.L;
  const long k1 = (k0 + a.L < a.n) ? k0 + a.L : a.n;
  const real cs = cscale ? *cscale : 1.0;
  LN::smooth(cx, k0, k1, ch == a.CS - 1, emit_t0 != 0, qinv, a.ql.v, sin + ch * ST, kern, cs, means, chols,
             part2 + ch * 2);
}

template <int d, int q>
struct Lane2Launchers {
  using LS = Lane2Setup<d, q>;
  template <class K>
  static cudaError_t prep(K kernel) {
    return ensure_smem(kernel, LS::smem_bytes());
  }
  static cudaError_t fold(cudaStream_t s, const LeafArgs& a, real* fagg, real* faggm) {
    if (cudaError_t e = prep(k_lane2_fold<d, q>)) return e;
    k_lane2_fold<d, q><<<LS::grid(a.CS), LANE2_WARPS * 32, LS::smem_bytes(), s>>>(a, fagg, faggm);
    return cudaGetLastError();
  }
  // the chunk smoothing elements come from the chunk-level op (pof_treelane.cuh): sagg must be null
  static cudaError_t scan(cudaStream_t s, const LeafArgs& a, const real* fin, real* kern, real* sagg,
                          real* send, real* part, real* fmeans, real* fchols) {
    if (sagg) return cudaErrorInvalidValue;
    if (cudaError_t e = ensure_smem(k_lane2_scan<d, q>, LS::smem_bytes(), true)) return e;
    k_lane2_scan<d, q><<<LS::grid(a.CS), LANE2_WARPS * 32, LS::smem_bytes(), s>>>(a, fin, kern, send, part, fmeans,
                                                                                  fchols);
    return cudaGetLastError();
  }
  static cudaError_t smooth(cudaStream_t s, const LeafArgs& a, const real* sin, const real* kern, int emit_t0,
                            const real* cscale, real* means, real* chols, real* part2) {
    if constexpr (LS::staged_ok()) {
      if (!a.no_tma) {
        if (cudaError_t e = ensure_smem(k_lane2_smooth<d, q, true>, LS::smem_bytes_staged())) return e;
        k_lane2_smooth<d, q, true><<<LS::grid(a.CS), LANE2_WARPS * 32, LS::smem_bytes_staged(), s>>>(
            a, sin, kern, emit_t0, cscale, means, chols, part2);
        return cudaGetLastError();
      }
    }
    if (cudaError_t e = prep(k_lane2_smooth<d, q, false>)) return e;
    k_lane2_smooth<d, q, false><<<LS::grid(a.CS), LANE2_WARPS * 32, LS::smem_bytes(), s>>>(a, sin, kern, emit_t0,
                                                                                           cscale, means, chols, part2);
    return cudaGetLastError();
  }
  static const LeafLaunch* get() {
    static const LeafLaunch l = {&fold, &scan, &smooth, nullptr, Lane2<d, q>::GPW, 1};
    return &l;
  }
};

}  // namespace POF_NS

// only state dimensions whose tree sweeps have the register-resident chunk-level op (2D <= 32)
#define POF_LANE2_CASE(dd, qq) \
  case qq:                     \
    if constexpr (dd * (qq + 1) <= 16) return Lane2Launchers<dd, qq>::get(); else return nullptr;
#define POF_DEFINE_LANE2_D(dd)                 \
  namespace POF_NS {                              \
  const LeafLaunch* lane2_launch_d##dd(int q) { \
    switch (q) {                               \
      POF_LANE2_CASE(dd, 1)                    \
      POF_LANE2_CASE(dd, 2)                    \
      POF_LANE2_CASE(dd, 3)                    \
      POF_LANE2_CASE(dd, 4)                    \
      POF_LANE2_CASE(dd, 5)                    \
      default: return nullptr;                 \
    }                                          \
  }                                            \
  }
