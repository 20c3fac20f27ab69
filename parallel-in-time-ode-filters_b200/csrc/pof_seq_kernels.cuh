// Sequential EKS (reference pof/sequential_filtsmooth/__init__.py:5-10, filter.py:9-30, smoother.py:8-28): extended
// Kalman filter relinearised at the PREDICTED mean of every step, then the RTS smoother.  Inherently sequential: one
// thread walks the whole grid with the plain-loop recursions of pof_leaf.cuh / pof_pipeline.cuh (the code the host
// simulator in tests/hostsim also compiles).  Baseline / cross-check path and the core of init="coarse"; d <= 4.
// Included by pof_seq_d{1,2,3,4}.cu (one translation unit per ODE dimension: parallel compilation).
#pragma once
#include "pof_ivp.cuh"
#include "pof_launch.cuh"
#include "pof_pipeline.cuh"

namespace pof {

template <int d, int q>
__global__ void __launch_bounds__(32)
    k_seq_eks(LeafArgs a, int ivp_id, IvpParams P, const double* __restrict__ x0, double* __restrict__ kern,
              double* __restrict__ means, double* __restrict__ chols, double* __restrict__ part) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Chunk<d, q>::seq_eks(a.n, a.s0, a.s1, a.ql.v, ivp_id, P, x0, kern, means, chols, part);
}

template <int d, int q>
struct SeqLaunchers {
  static cudaError_t seq_eks(cudaStream_t s, const LeafArgs& a, int ivp_id, const double* params8, const double* x0,
                             double* kern, double* means, double* chols, double* part) {
    IvpParams P;
    for (int i = 0; i < 8; ++i) P.p[i] = params8[i];
    k_seq_eks<d, q><<<1, 32, 0, s>>>(a, ivp_id, P, x0, kern, means, chols, part);
    return cudaGetLastError();
  }
  static const LeafLaunch* get() {
    static const LeafLaunch l = {nullptr, nullptr, nullptr, &seq_eks, 0, 0, 0};
    return &l;
  }
};

}  // namespace pof

#define POF_DEFINE_SEQ_D(dd)                                   \
  namespace pof {                                              \
  const LeafLaunch* seq_launch_d##dd(int q) {                  \
    switch (q) {                                               \
      case 1: return SeqLaunchers<dd, 1>::get();               \
      case 2: return SeqLaunchers<dd, 2>::get();               \
      case 3: return SeqLaunchers<dd, 3>::get();               \
      case 4: return SeqLaunchers<dd, 4>::get();               \
      case 5: return SeqLaunchers<dd, 5>::get();               \
      default: return nullptr;                                 \
    }                                                          \
  }                                                            \
  }
