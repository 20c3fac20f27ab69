// Tree-sweep kernels over chunk carries with the register-resident combines of pof_treelane.cuh.
// One group of lanes per tree node; `which` selects the sweep.
#pragma once
#include "pof_launch.cuh"
#include "pof_treelane.cuh"

namespace pof {

constexpr int TL_WARPS = 4;

enum TreeOp { T_FUP = 0, T_FDOWN, T_SUP, T_SDOWN, T_FCOMB, T_SCOMB, T_CHUNKK };

template <int D, int G>
__device__ __forceinline__ void group_copy(int r, double* __restrict__ dst, const double* __restrict__ src, int n) {
  for (int i = r; i < n; i += G) dst[i] = src[i];
}

// a, b, c, na, nb follow the generic kernels in pof_api.cu:
//  T_FUP   : a = children elems, na = #children, c = parent elems, nb = #parents
//  T_FDOWN : a = parent states, nb = #parents, b = children elems, na = #children, c = children states
//  T_SUP / T_SDOWN: same with smoothing elements
//  T_FCOMB / T_SCOMB: c[i] = op(a[i], b[i]), nb = count
template <int D, int OP>
__global__ void __launch_bounds__(TL_WARPS * 32)
    k_tree(const double* __restrict__ a, long na, const double* __restrict__ b, double* __restrict__ c, long nb) {
  extern __shared__ __align__(16) double sm[];
  using TL = TreeLane<D>;
  constexpr bool FILT = (OP == T_FUP || OP == T_FDOWN || OP == T_FCOMB || OP == T_CHUNKK);
  constexpr int G = FILT ? TL::G2 : TL::GS;
  constexpr int CPW = 32 / G;
  constexpr int FE = 3 * D * D + 2 * D, SE = 2 * D * D + D, ST = D * D + D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long i = ((long)blockIdx.x * TL_WARPS + warp) * CPW + lane / G;
  if (i >= nb) return;
  typename TL::Ctx cx;
  TL::template init<G>(cx, sm + (warp * CPW + lane / G) * TL::SM_COMBINE);
  if constexpr (OP == T_FUP) {
    const double* lc = a + 2 * i * FE;
    if (2 * i + 1 < na) TL::template filter_combine<false>(cx, lc, lc + FE, c + i * FE);
    else group_copy<D, G>(cx.r, c + i * FE, lc, FE);
  } else if constexpr (OP == T_FDOWN) {
    const double* p = a + i * ST;
    group_copy<D, G>(cx.r, c + 2 * i * ST, p, ST);
    if (2 * i + 1 < na) TL::template filter_combine<true>(cx, p, b + 2 * i * FE, c + (2 * i + 1) * ST);
  } else if constexpr (OP == T_SUP) {
    const double* lc = a + 2 * i * SE;
    if (2 * i + 1 < na) TL::template smooth_combine<false>(cx, lc + SE, lc, c + i * SE);
    else group_copy<D, G>(cx.r, c + i * SE, lc, SE);
  } else if constexpr (OP == T_SDOWN) {
    const double* p = a + i * ST;
    if (2 * i + 1 < na) {
      group_copy<D, G>(cx.r, c + (2 * i + 1) * ST, p, ST);
      TL::template smooth_combine<true>(cx, p, b + (2 * i + 1) * SE, c + 2 * i * ST);
    } else {
      group_copy<D, G>(cx.r, c + 2 * i * ST, p, ST);
    }
  } else if constexpr (OP == T_CHUNKK) {
    // a = chunk incoming states, b = chunk filtering elements before their last update, c = chunk smoothing elements
    TL::chunk_kernel(cx, a + i * ST, b + i * FE, c + i * SE);
  } else if constexpr (OP == T_FCOMB) {
    TL::template filter_combine<false>(cx, a + i * FE, b + i * FE, c + i * FE);
  } else {
    TL::template smooth_combine<false>(cx, a + i * SE, b + i * SE, c + i * SE);
  }
}

template <int D>
struct TreeLaunchers {
  using TL = TreeLane<D>;
  template <int OP>
  static cudaError_t run(cudaStream_t s, const double* a, long na, const double* b, double* c, long nb) {
    constexpr bool FILT = (OP == T_FUP || OP == T_FDOWN || OP == T_FCOMB || OP == T_CHUNKK);
    constexpr int G = FILT ? TL::G2 : TL::GS;
    constexpr int CPW = 32 / G;
    constexpr int smem = TL_WARPS * CPW * TL::SM_COMBINE * (int)sizeof(double);
    if (nb <= 0) return cudaSuccess;
    if (cudaError_t e = ensure_smem(k_tree<D, OP>, smem)) return e;
    const long per_block = (long)TL_WARPS * CPW;
    k_tree<D, OP><<<(unsigned)((nb + per_block - 1) / per_block), TL_WARPS * 32, smem, s>>>(a, na, b, c, nb);
    return cudaGetLastError();
  }
  static const TreeLaunch* get() {
    static const TreeLaunch t = {&run<T_FUP>, &run<T_FDOWN>, &run<T_SUP>, &run<T_SDOWN>, &run<T_FCOMB>, &run<T_SCOMB>,
                                 &run<T_CHUNKK>};
    return &t;
  }
};

}  // namespace pof
