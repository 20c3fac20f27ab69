// Leaf kernels: one thread per time-chunk (see pof_pipeline.cuh / pof_leaf.cuh for the math).
// Included by pof_leaf_d{1,2,3,4}.cu with POF_LEAF_D defined.
#pragma once
#include "pof_ivp.cuh"
#include "pof_launch.cuh"
#include "pof_pipeline.cuh"

namespace pof {

constexpr int LEAF_THREADS = 128;

template <int d, int q>
__global__ void __launch_bounds__(LEAF_THREADS) k_fold(LeafArgs a, double* __restrict__ fagg) {
  const long ch = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= a.CS) return;
  constexpr int D = Chunk<d, q>::D;
  constexpr int FE = 3 * D * D + 2 * D;
  const long k0 = ch * a.L;
  const long k1 = (k0 + a.L < a.n) ? k0 + a.L : a.n;
  Chunk<d, q>::fold(k0, k1, a.H, a.c, a.ql.v, fagg + ch * FE);
}

template <int d, int q>
__global__ void __launch_bounds__(LEAF_THREADS)
    k_scan(LeafArgs a, const double* __restrict__ fin, double* __restrict__ kern, double* __restrict__ sagg,
           double* __restrict__ send, double* __restrict__ part, double* __restrict__ fmeans,
           double* __restrict__ fchols) {
  const long ch = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= a.CS) return;
  constexpr int D = Chunk<d, q>::D;
  constexpr int SE = 2 * D * D + D, ST = D * D + D;
  const long k0 = ch * a.L;
  const long k1 = (k0 + a.L < a.n) ? k0 + a.L : a.n;
  Chunk<d, q>::scan(k0, k1, a.H, a.c, a.ql.v, fin + ch * ST, kern, a.CS, ch, sagg + ch * SE, send + ch * ST,
                    part + ch * 3, fmeans, fchols);
}

template <int d, int q>
__global__ void __launch_bounds__(LEAF_THREADS)
    k_smooth(LeafArgs a, const double* __restrict__ sin, const double* __restrict__ kern, int emit_t0,
             const double* __restrict__ cscale, double* __restrict__ means, double* __restrict__ chols,
             double* __restrict__ part2) {
  const long ch = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= a.CS) return;
  constexpr int D = Chunk<d, q>::D;
  constexpr int ST = D * D + D;
  const long k0 = ch * a.L;
  const long k1 = (k0 + a.L < a.n) ? k0 + a.L : a.n;
  const double cs = cscale ? *cscale : 1.0;
  Chunk<d, q>::smooth(k0, k1, ch == a.CS - 1, emit_t0 != 0, a.ql.v, sin + ch * ST, kern, a.CS, ch, cs, means, chols,
                      part2 + ch * 2);
}

// Sequential EKS (reference pof/sequential_filtsmooth/__init__.py:5-10, filter.py:9-30, smoother.py:8-28): extended
// Kalman filter relinearised at the PREDICTED mean of every step, then the RTS smoother.  Inherently sequential:
// one thread walks the whole grid (baseline / cross-check path, not a performance path).
template <int d, int q>
__global__ void __launch_bounds__(32)
    k_seq_eks(LeafArgs a, int ivp_id, IvpParams P, const double* __restrict__ x0, double* __restrict__ kern,
              double* __restrict__ means, double* __restrict__ chols, double* __restrict__ part) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Chunk<d, q>::seq_eks(a.n, a.s0, a.s1, a.ql.v, ivp_id, P, x0, kern, means, chols, part);
}

template <int d, int q>
struct LeafLaunchers {
  static cudaError_t seq_eks(cudaStream_t s, const LeafArgs& a, int ivp_id, const double* params8, const double* x0,
                             double* kern, double* means, double* chols, double* part) {
    IvpParams P;
    for (int i = 0; i < 8; ++i) P.p[i] = params8[i];
    k_seq_eks<d, q><<<1, 32, 0, s>>>(a, ivp_id, P, x0, kern, means, chols, part);
    return cudaGetLastError();
  }
  static unsigned grid(const LeafArgs& a) { return (unsigned)((a.CS + LEAF_THREADS - 1) / LEAF_THREADS); }
  static cudaError_t fold(cudaStream_t s, const LeafArgs& a, double* fagg, double* /*faggm*/) {
    k_fold<d, q><<<grid(a), LEAF_THREADS, 0, s>>>(a, fagg);
    return cudaGetLastError();
  }
  static cudaError_t scan(cudaStream_t s, const LeafArgs& a, const double* fin, double* kern, double* sagg,
                          double* send, double* part, double* fmeans, double* fchols) {
    k_scan<d, q><<<grid(a), LEAF_THREADS, 0, s>>>(a, fin, kern, sagg, send, part, fmeans, fchols);
    return cudaGetLastError();
  }
  static cudaError_t smooth(cudaStream_t s, const LeafArgs& a, const double* sin, const double* kern, int emit_t0,
                            const double* cscale, double* means, double* chols, double* part2) {
    k_smooth<d, q><<<grid(a), LEAF_THREADS, 0, s>>>(a, sin, kern, emit_t0, cscale, means, chols, part2);
    return cudaGetLastError();
  }
  static const LeafLaunch* get() {
    static const LeafLaunch l = {&fold, &scan, &smooth, &seq_eks, 32, 0};
    return &l;
  }
};

}  // namespace pof

#define POF_DEFINE_LEAF_D(dd)                                   \
  namespace pof {                                               \
  const LeafLaunch* leaf_launch_d##dd(int q) {                  \
    switch (q) {                                                \
      case 1: return LeafLaunchers<dd, 1>::get();               \
      case 2: return LeafLaunchers<dd, 2>::get();               \
      case 3: return LeafLaunchers<dd, 3>::get();               \
      case 4: return LeafLaunchers<dd, 4>::get();               \
      case 5: return LeafLaunchers<dd, 5>::get();               \
      default: return nullptr;                                  \
    }                                                           \
  }                                                             \
  }
