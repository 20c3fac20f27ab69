// Tree-sweep kernels over chunk carries with the register-resident combines of pof_treelane.cuh.
// One group of lanes per tree node; `which` selects the sweep.
#pragma once
#include <cooperative_groups.h>

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

// ---------------------------------------------------------------------------------------------------------------
// Whole sweep in one cooperative kernel.  Every group of G lanes walks the nodes of the current level with a grid
// stride; levels are separated by grid barriers (all CTAs are co-resident: cooperative launch).
template <int D, bool FILT>
__global__ void __launch_bounds__(TL_WARPS * 32) k_tree_sweep(const SweepArgs A) {
  extern __shared__ __align__(16) double sm[];
  namespace cg = cooperative_groups;
  cg::grid_group grid = cg::this_grid();
  using TL = TreeLane<D>;
  constexpr int G = FILT ? TL::G2 : TL::GS;
  constexpr int G2 = TL::G2;  // the chunk-level op always uses the filtering group width
  constexpr int FE = 3 * D * D + 2 * D, SE = 2 * D * D + D, ST = D * D + D;
  constexpr int EL = FILT ? FE : SE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if constexpr (!FILT) {
    if (A.faggm) {  // chunk-level smoothing elements (pof_treelane.cuh: chunk_kernel), G2 lanes per chunk
      constexpr int CPW2 = 32 / G2;
      typename TL::Ctx cx;
      TL::template init<G2>(cx, sm + (warp * CPW2 + lane / G2) * TL::SM_COMBINE);
      const long ng = (long)gridDim.x * TL_WARPS * CPW2;
      for (long i = ((long)blockIdx.x * TL_WARPS + warp) * CPW2 + lane / G2; i < A.sz[0]; i += ng)
        TL::chunk_kernel(cx, A.fin + i * ST, A.faggm + i * FE, A.agg + i * SE);
      if (A.block_sync) __syncthreads(); else grid.sync();
    }
  }
  constexpr int CPW = 32 / G;
  typename TL::Ctx cx;
  TL::template init<G>(cx, sm + (warp * CPW + lane / G) * TL::SM_COMBINE);
  const long g0 = ((long)blockIdx.x * TL_WARPS + warp) * CPW + lane / G;
  const long ng = (long)gridDim.x * TL_WARPS * CPW;
  // ---- up-sweep
  for (int l = A.up_begin; l < A.up_end; ++l) {
    const double* ch = A.agg + A.off[l] * EL;
    double* pa = A.agg + A.off[l + 1] * EL;
    const long na = A.sz[l], nb = A.sz[l + 1];
    for (long i = g0; i < nb; i += ng) {
      const double* lc = ch + 2 * i * EL;
      if (2 * i + 1 < na) {
        if constexpr (FILT) TL::template filter_combine<false>(cx, lc, lc + EL, pa + i * EL);
        else TL::template smooth_combine<false>(cx, lc + EL, lc, pa + i * EL);
      } else {
        group_copy<D, G>(cx.r, pa + i * EL, lc, EL);
      }
    }
    if (A.block_sync) __syncthreads(); else grid.sync();
  }
  if (A.down_begin <= A.down_end) return;
  // ---- root state
  if (A.root_m && g0 == 0) {
    double* r = A.st + A.off[A.nlev - 1] * ST;
    for (int i = cx.r; i < ST; i += G) r[i] = (i < D) ? A.root_m[i] : A.root_L[i - D];
  }
  if (A.root_m) {
    if (A.block_sync) __syncthreads(); else grid.sync();
  }
  // ---- down-sweep (state form)
  for (int l = A.down_begin; l > A.down_end; --l) {
    const double* ps = A.st + A.off[l] * ST;
    const double* el = A.agg + A.off[l - 1] * EL;
    double* cs = A.st + A.off[l - 1] * ST;
    const long na = A.sz[l - 1], nb = A.sz[l];
    for (long i = g0; i < nb; i += ng) {
      const double* p = ps + i * ST;
      if constexpr (FILT) {
        group_copy<D, G>(cx.r, cs + 2 * i * ST, p, ST);
        if (2 * i + 1 < na) TL::template filter_combine<true>(cx, p, el + 2 * i * EL, cs + (2 * i + 1) * ST);
      } else {
        if (2 * i + 1 < na) {
          group_copy<D, G>(cx.r, cs + (2 * i + 1) * ST, p, ST);
          TL::template smooth_combine<true>(cx, p, el + (2 * i + 1) * EL, cs + 2 * i * ST);
        } else {
          group_copy<D, G>(cx.r, cs + 2 * i * ST, p, ST);
        }
      }
    }
    if (l > A.down_end + 1) {
      if (A.block_sync) __syncthreads(); else grid.sync();
    }
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
    if (cudaError_t e = ensure_smem(k_tree<D, OP>, smem, OP == T_SUP)) return e;
    const long per_block = (long)TL_WARPS * CPW;
    k_tree<D, OP><<<(unsigned)((nb + per_block - 1) / per_block), TL_WARPS * 32, smem, s>>>(a, na, b, c, nb);
    return cudaGetLastError();
  }
  // cooperative whole-sweep launch; the grid is the smaller of "all CTAs co-resident" and "one group per node of the
  // widest level"
  template <bool FILT>
  static cudaError_t sweep(cudaStream_t s, const SweepArgs& A) {
    constexpr int G = FILT ? TL::G2 : TL::GS;
    constexpr int CPWmin = 32 / TL::G2;  // the chunk-level prologue of the smoothing sweep uses G2 lanes per item
    constexpr int smem = TL_WARPS * (32 / G) * TL::SM_COMBINE * (int)sizeof(double);
    auto kern = k_tree_sweep<D, FILT>;
    if (cudaError_t e = ensure_smem(kern, smem)) return e;
    static int max_grid[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (max_grid[dev] == 0) {
      int per_sm = 0, sms = 0;
      if (cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TL_WARPS * 32, smem)) return e;
      if (cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) return e;
      if (per_sm < 1) return cudaErrorLaunchOutOfResources;
      max_grid[dev] = per_sm * sms;
    }
    if (A.block_sync) {  // the apex: one CTA, plain launch
      k_tree_sweep<D, FILT><<<1, TL_WARPS * 32, smem, s>>>(A);
      return cudaGetLastError();
    }
    long widest = 1;
    if (A.up_end > A.up_begin) widest = A.sz[A.up_begin + 1];
    if (A.down_begin > A.down_end && A.sz[A.down_end + 1] > widest) widest = A.sz[A.down_end + 1];
    long blocks = (widest + (long)TL_WARPS * (32 / G) - 1) / ((long)TL_WARPS * (32 / G));
    if (!FILT && A.faggm) {
      const long b2 = (A.sz[0] + (long)TL_WARPS * CPWmin - 1) / ((long)TL_WARPS * CPWmin);
      if (b2 > blocks) blocks = b2;
    }
    if (blocks < 1) blocks = 1;
    if (blocks > max_grid[dev]) blocks = max_grid[dev];
    SweepArgs a = A;
    void* params[] = {(void*)&a};
    return cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)blocks), dim3(TL_WARPS * 32), params,
                                       (size_t)smem, s);
  }
  static const TreeLaunch* get() {
    static const TreeLaunch t = {&run<T_FUP>, &run<T_FDOWN>, &run<T_SUP>, &run<T_SDOWN>, &run<T_FCOMB>, &run<T_SCOMB>,
                                 &run<T_CHUNKK>, &sweep<true>, &sweep<false>,
                                 TL_WARPS * (32 / TL::G2), TL_WARPS * (32 / TL::GS)};
    return &t;
  }
};

}  // namespace pof
