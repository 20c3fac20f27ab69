// Tree-sweep kernels over chunk carries with the register-resident combines of pof_treelane.cuh.
// One group of lanes per tree node; `which` selects the sweep.
#pragma once
#include "pof_launch.cuh"
#include "pof_treelane.cuh"

namespace pof {

constexpr int TL_WARPS = 4;

enum TreeOp { T_FUP = 0, T_FDOWN, T_SUP, T_SDOWN, T_FCOMB, T_SCOMB, T_CHUNKK };

// (L2 loads: inside a dataflow sweep the source was written by another SM during the SAME kernel)
template <int D, int G>
__device__ __forceinline__ void group_copy(int r, double* dst, const double* src, int n) {
  for (int i = r; i < n; i += G) dst[i] = __ldcg(src + i);
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
// Whole sweep as one dataflow kernel (FlowArgs, pof_launch.cuh): tickets in level order, per-node ready flags.
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int D, bool FILT>
__global__ void __launch_bounds__(TL_WARPS * 32) k_tree_flow(const FlowArgs A) {
  extern __shared__ __align__(16) double sm[];
  using TL = TreeLane<D>;
  constexpr int G = FILT ? TL::G2 : TL::GS;
  constexpr int CPW = 32 / G;
  constexpr int FE = 3 * D * D + 2 * D, SE = 2 * D * D + D, ST = D * D + D;
  constexpr int EL = FILT ? FE : SE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  typename TL::Ctx cx;
  TL::template init<G>(cx, sm + (warp * CPW + lane / G) * TL::SM_COMBINE);
  const long total = A.seg_begin[A.nseg];
  while (true) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(A.ticket, 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    const long first = (long)t * CPW;
    if (first >= total) break;
    const long item = first + lane / G;
    // decode: segment, level, index within the segment (segments are padded to whole tickets: dead items at the end)
    int kind = FlowArgs::ROOT, lev = 0;
    long i = 0;
    bool live = false;
    {
      int sgm = 0;
      while (sgm + 1 < A.nseg && item >= A.seg_begin[sgm + 1]) ++sgm;
      kind = A.seg_kind[sgm];
      lev = A.seg_level[sgm];
      i = item - A.seg_begin[sgm];
      live = i < A.seg_count[sgm];
    }
    // dependencies (at most two flags)
    const unsigned* dep0 = nullptr;
    const unsigned* dep1 = nullptr;
    if (live && kind == FlowArgs::UP) {
      if (lev - 1 >= A.up_lo && lev - 1 <= A.up_hi) {
        dep0 = A.flag_up + A.off[lev - 1] + 2 * i;
        if (2 * i + 1 < A.sz[lev - 1]) dep1 = dep0 + 1;
      }
    } else if (live && kind == FlowArgs::DOWN) {
      dep0 = A.flag_dn + A.off[lev] + i;
      const long e = FILT ? 2 * i : 2 * i + 1;  // the child element the combine reads
      if (2 * i + 1 < A.sz[lev - 1] && lev - 1 >= A.up_lo && lev - 1 <= A.up_hi) dep1 = A.flag_up + A.off[lev - 1] + e;
    }
    while (true) {
      const bool ok = (!dep0 || ld_acquire(dep0) != 0u) && (!dep1 || ld_acquire(dep1) != 0u);
      if (__all_sync(0xffffffffu, ok)) break;
      __nanosleep(64);
    }
    __syncwarp();
    if (live) {
      if (kind == FlowArgs::UP) {
        const double* lc = A.agg + (A.off[lev - 1] + 2 * i) * EL;
        double* pa = A.agg + (A.off[lev] + i) * EL;
        if (2 * i + 1 < A.sz[lev - 1]) {
          if constexpr (FILT) TL::template filter_combine<false>(cx, lc, lc + EL, pa);
          else TL::template smooth_combine<false>(cx, lc + EL, lc, pa);
        } else {
          group_copy<D, G>(cx.r, pa, lc, EL);
        }
        __threadfence();
        cx.sync();
        if (cx.r == 0) st_release(A.flag_up + A.off[lev] + i, 1u);
      } else if (kind == FlowArgs::ROOT) {
        double* r = A.st + A.off[A.nlev - 1] * ST;
        for (int j = cx.r; j < ST; j += G) r[j] = (j < D) ? __ldcg(A.root_m + j) : __ldcg(A.root_L + (j - D));
        __threadfence();
        cx.sync();
        if (cx.r == 0) st_release(A.flag_dn + A.off[A.nlev - 1], 1u);
      } else {
        const double* p = A.st + (A.off[lev] + i) * ST;
        const double* el = A.agg + A.off[lev - 1] * EL;
        double* cs = A.st + A.off[lev - 1] * ST;
        const bool two = 2 * i + 1 < A.sz[lev - 1];
        if constexpr (FILT) {
          group_copy<D, G>(cx.r, cs + 2 * i * ST, p, ST);
          if (two) TL::template filter_combine<true>(cx, p, el + 2 * i * EL, cs + (2 * i + 1) * ST);
        } else {
          if (two) {
            group_copy<D, G>(cx.r, cs + (2 * i + 1) * ST, p, ST);
            TL::template smooth_combine<true>(cx, p, el + (2 * i + 1) * EL, cs + 2 * i * ST);
          } else {
            group_copy<D, G>(cx.r, cs + 2 * i * ST, p, ST);
          }
        }
        __threadfence();
        cx.sync();
        if (cx.r == 0) {
          st_release(A.flag_dn + A.off[lev - 1] + 2 * i, 1u);
          if (two) st_release(A.flag_dn + A.off[lev - 1] + 2 * i + 1, 1u);
        }
      }
    }
    __syncwarp();
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
  // dataflow whole-sweep launch: persistent grid (enough warps for the widest level, at most what is resident)
  template <bool FILT>
  static cudaError_t flow(cudaStream_t s, const FlowArgs& A) {
    constexpr int G = FILT ? TL::G2 : TL::GS;
    constexpr int CPW = 32 / G;
    constexpr int smem = TL_WARPS * CPW * TL::SM_COMBINE * (int)sizeof(double);
    auto kern = k_tree_flow<D, FILT>;
    if (cudaError_t e = ensure_smem(kern, smem, !FILT)) return e;
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
    FlowArgs a = A;
    long widest = 1;
    a.seg_begin[0] = 0;
    for (int j = 0; j < a.nseg; ++j) {
      const long n = a.seg_count[j];
      if (n > widest) widest = n;
      a.seg_begin[j + 1] = a.seg_begin[j] + (n + CPW - 1) / CPW * CPW;
    }
    long blocks = (widest + (long)TL_WARPS * CPW - 1) / ((long)TL_WARPS * CPW);
    if (blocks < 1) blocks = 1;
    if (blocks > max_grid[dev]) blocks = max_grid[dev];
    k_tree_flow<D, FILT><<<(unsigned)blocks, TL_WARPS * 32, smem, s>>>(a);
    return cudaGetLastError();
  }
  static const TreeLaunch* get() {
    static const TreeLaunch t = {&run<T_FUP>, &run<T_FDOWN>, &run<T_SUP>, &run<T_SDOWN>, &run<T_FCOMB>, &run<T_SCOMB>,
                                 &run<T_CHUNKK>, &flow<true>, &flow<false>};
    return &t;
  }
};

}  // namespace pof
