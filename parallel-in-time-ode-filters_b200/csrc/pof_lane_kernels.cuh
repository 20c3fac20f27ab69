// Leaf kernels, performance path: one group of G lanes per time-chunk (pof_lane.cuh).
#pragma once
#include "pof_lane.cuh"
#include "pof_launch.cuh"

namespace pof {

constexpr int LANE_WARPS = 4;
// resident CTAs per SM the compiler must allow.  Measured on B200: forcing more occupancy (168 / 128 registers) makes
// fold and smooth SLOWER (spills add LSU traffic, and the kernels are bound by shared-memory delivery, not by warps)
#ifndef POF_FOLD_MINBLOCKS
#define POF_FOLD_MINBLOCKS 1
#endif
#ifndef POF_SCAN_MINBLOCKS
#define POF_SCAN_MINBLOCKS 1
#endif
#ifndef POF_SMOOTH_MINBLOCKS
#define POF_SMOOTH_MINBLOCKS 2
#endif

template <int d, int q>
struct LaneSetup {
  using LN = Lane<d, q>;
  static __device__ __forceinline__ long chunk_of_thread() {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return ((long)blockIdx.x * LANE_WARPS + warp) * LN::GPW + lane / LN::G;
  }
  static __device__ __forceinline__ double* smem_of_thread(double* sm) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    return sm + (warp * LN::GPW + lane / LN::G) * LN::SM_GROUP;
  }
  static constexpr int smem_bytes() { return LANE_WARPS * LN::GPW * LN::SM_GROUP * (int)sizeof(double); }
  static unsigned grid(long CS) {
    const long per_block = (long)LANE_WARPS * LN::GPW;
    return (unsigned)((CS + per_block - 1) / per_block);
  }
};

template <int d, int q>
__global__ void __launch_bounds__(LANE_WARPS * 32, POF_FOLD_MINBLOCKS)
    k_lane_fold(LeafArgs a, double* __restrict__ fagg, double* __restrict__ faggm) {
  extern __shared__ __align__(16) double sm[];
  using LN = Lane<d, q>;
  const long ch = LaneSetup<d, q>::chunk_of_thread();
  if (ch >= a.CS) return;
  typename LN::Ctx cx;
  LN::init_ctx(cx, LaneSetup<d, q>::smem_of_thread(sm), a.ql.v);
  const long k0 = ch * a.L;
  const long k1 = (k0 + a.L < a.n) ? k0 + a.L : a.n;
  constexpr int FE = 3 * LN::D * LN::D + 2 * LN::D;
  const typename LN::Lin lin = {a.H, a.c, a.Jc, a.s0, a.s1};
  LN::fold(cx, k0, k1, lin, fagg + ch * FE, faggm ? faggm + ch * FE : nullptr);
}

template <int d, int q, bool COMPOSE>
__global__ void __launch_bounds__(LANE_WARPS * 32, POF_SCAN_MINBLOCKS)
    k_lane_scan(LeafArgs a, const double* __restrict__ fin, double* __restrict__ kern, double* __restrict__ sagg,
                double* __restrict__ send, double* __restrict__ part, double* __restrict__ fmeans,
                double* __restrict__ fchols) {
  extern __shared__ __align__(16) double sm[];
  using LN = Lane<d, q>;
  const long ch = LaneSetup<d, q>::chunk_of_thread();
  if (ch >= a.CS) return;
  typename LN::Ctx cx;
  LN::init_ctx(cx, LaneSetup<d, q>::smem_of_thread(sm), a.ql.v);
  constexpr int D = LN::D, SE = 2 * D * D + D, ST = D * D + D;
  const long k0 = ch * a.L;
  const long k1 = (k0 + a.L < a.n) ? k0 + a.L : a.n;
  const typename LN::Lin lin = {a.H, a.c, a.Jc, a.s0, a.s1};
  LN::template scan<COMPOSE>(cx, k0, k1, lin, fin + ch * ST, kern, COMPOSE ? sagg + ch * SE : nullptr,
                             send + ch * ST, part + ch * 3, fmeans, fchols);
}

template <int d, int q>
__global__ void __launch_bounds__(LANE_WARPS * 32, POF_SMOOTH_MINBLOCKS)
    k_lane_smooth(LeafArgs a, const double* __restrict__ sin, const double* __restrict__ kern, int emit_t0,
                  const double* __restrict__ cscale, double* __restrict__ means, double* __restrict__ chols,
                  double* __restrict__ part2) {
  extern __shared__ __align__(16) double sm[];
  using LN = Lane<d, q>;
  const long ch = LaneSetup<d, q>::chunk_of_thread();
  if (ch >= a.CS) return;
  typename LN::Ctx cx;
  LN::init_ctx(cx, LaneSetup<d, q>::smem_of_thread(sm), a.ql.v);
  constexpr int D = LN::D, ST = D * D + D;
  double qinv[LN::Q1];
#pragma unroll
  for (int i = 0; i < LN::Q1; ++i) qinv[i] = 1.0 / a.ql.v[i * LN::Q1 + i];
  const long k0 = ch * a.L;
  const long k1 = (k0 + a.L < a.n) ? k0 + a.L : a.n;
  const double cs = cscale ? *cscale : 1.0;
  LN::smooth(cx, k0, k1, ch == a.CS - 1, emit_t0 != 0, qinv, a.ql.v, sin + ch * ST, kern, cs, means, chols,
             part2 + ch * 2);
}

template <int d, int q>
struct LaneLaunchers {
  using LS = LaneSetup<d, q>;
  static cudaError_t fold(cudaStream_t s, const LeafArgs& a, double* fagg, double* faggm) {
    k_lane_fold<d, q><<<LS::grid(a.CS), LANE_WARPS * 32, LS::smem_bytes(), s>>>(a, fagg, faggm);
    return cudaGetLastError();
  }
  static cudaError_t scan(cudaStream_t s, const LeafArgs& a, const double* fin, double* kern, double* sagg,
                          double* send, double* part, double* fmeans, double* fchols) {
    if (sagg)
      k_lane_scan<d, q, true><<<LS::grid(a.CS), LANE_WARPS * 32, LS::smem_bytes(), s>>>(a, fin, kern, sagg, send,
                                                                                         part, fmeans, fchols);
    else
      k_lane_scan<d, q, false><<<LS::grid(a.CS), LANE_WARPS * 32, LS::smem_bytes(), s>>>(a, fin, kern, sagg, send,
                                                                                          part, fmeans, fchols);
    return cudaGetLastError();
  }
  static cudaError_t smooth(cudaStream_t s, const LeafArgs& a, const double* sin, const double* kern, int emit_t0,
                            const double* cscale, double* means, double* chols, double* part2) {
    k_lane_smooth<d, q><<<LS::grid(a.CS), LANE_WARPS * 32, LS::smem_bytes(), s>>>(a, sin, kern, emit_t0, cscale,
                                                                                   means, chols, part2);
    return cudaGetLastError();
  }
  static const LeafLaunch* get() {
    static const LeafLaunch l = {&fold, &scan, &smooth, nullptr, Lane<d, q>::GPW, 1};
    return &l;
  }
};

}  // namespace pof

#define POF_DEFINE_LANE_D(dd)                                   \
  namespace pof {                                               \
  const LeafLaunch* lane_launch_d##dd(int q) {                  \
    switch (q) {                                                \
      case 1: return LaneLaunchers<dd, 1>::get();               \
      case 2: return LaneLaunchers<dd, 2>::get();               \
      case 3: return LaneLaunchers<dd, 3>::get();               \
      case 4: return LaneLaunchers<dd, 4>::get();               \
      case 5: return LaneLaunchers<dd, 5>::get();               \
      default: return nullptr;                                  \
    }                                                           \
  }                                                             \
  }
