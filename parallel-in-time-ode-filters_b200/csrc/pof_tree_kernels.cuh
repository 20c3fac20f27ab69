// Tree-sweep kernels over chunk carries with the register-resident combines of pof_treelane.cuh.
// One group of lanes per tree node; `which` selects the sweep.
#pragma once
#include "pof_launch.cuh"
#include "pof_treelane.cuh"

namespace POF_NS {

constexpr int TL_WARPS = 4;

enum TreeOp { T_FUP = 0, T_FDOWN, T_SUP, T_SDOWN, T_FCOMB, T_SCOMB, T_CHUNKK, T_SSEED };

// (L2 loads: inside a dataflow sweep the source was written by another SM during the SAME kernel)
template <int D, int G>
__device__ __forceinline__ void group_copy(int r, real* dst, const real* src, int n) {
  for (int i = r; i < n; i += G) dst[i] = __ldcg(src + i);
}

// a, b, c, na, nb follow the generic kernels in pof_api.cu:
//  T_FUP   : a = children elems, na = #children, c = parent elems, nb = #parents
//  T_FDOWN : a = parent states, nb = #parents, b = children elems, na = #children, c = children states
//  T_SUP / T_SDOWN: same with smoothing elements
//  T_FCOMB / T_SCOMB: c[i] = op(a[i], b[i]), nb = count
//  T_CHUNKK: a = chunk incoming states, b = chunk filtering elements before their last update, c = chunk smoothing elems
//  T_SSEED : a = ONE terminal state (D + D^2), b = per-chunk exclusive-suffix smoothing elements (the aggregate of all
//            LATER chunks), c = per-chunk seeds: c[i] = state-form op(a, b[i])   (nb chunks)
// WARPS: warps per CTA (2 for the kernels that run on the side stream next to the resident filter-scan CTAs: with
// two warps any register count fits into what those leave free)
template <int D, int OP, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
    k_tree(const real* __restrict__ a, long na, const real* __restrict__ b, real* __restrict__ c, long nb) {
  extern __shared__ __align__(16) real sm[];
  using TL = TreeLane<D>;
  constexpr bool FILT = (OP == T_FUP || OP == T_FDOWN || OP == T_FCOMB || OP == T_CHUNKK);
  constexpr int G = FILT ? TL::G2 : TL::GS;
  constexpr int CPW = 32 / G;
  constexpr int SMC = FILT ? TL::SM_COMBINE : TL::SM_COMBINE_S;
  constexpr int FE = 3 * D * D + 2 * D, SE = 2 * D * D + D, ST = D * D + D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long i = ((long)blockIdx.x * WARPS + warp) * CPW + lane / G;
  if (i >= nb) return;
  typename TL::Ctx cx;
  TL::template init<G>(cx, sm + (warp * CPW + lane / G) * SMC, FILT ? TL::NMAT : TL::NMAT_S);
  if constexpr (OP == T_FUP) {
    const real* lc = a + 2 * i * FE;
    if (2 * i + 1 < na) TL::template filter_combine<false>(cx, lc, lc + FE, c + i * FE);
    else group_copy<D, G>(cx.r, c + i * FE, lc, FE);
  } else if constexpr (OP == T_FDOWN) {
    const real* p = a + i * ST;
    group_copy<D, G>(cx.r, c + 2 * i * ST, p, ST);
    if (2 * i + 1 < na) TL::template filter_combine<true>(cx, p, b + 2 * i * FE, c + (2 * i + 1) * ST);
  } else if constexpr (OP == T_SUP) {
    const real* lc = a + 2 * i * SE;
    if (2 * i + 1 < na) TL::template smooth_combine<false>(cx, lc + SE, lc, c + i * SE);
    else group_copy<D, G>(cx.r, c + i * SE, lc, SE);
  } else if constexpr (OP == T_SDOWN) {
    const real* p = a + i * ST;
    if (2 * i + 1 < na) {
      group_copy<D, G>(cx.r, c + (2 * i + 1) * ST, p, ST);
      TL::template smooth_combine<true>(cx, p, b + (2 * i + 1) * SE, c + 2 * i * ST);
    } else {
      group_copy<D, G>(cx.r, c + 2 * i * ST, p, ST);
    }
  } else if constexpr (OP == T_CHUNKK) {
    // a = chunk incoming states, b = chunk filtering elements before their last update, c = chunk smoothing elements
    TL::chunk_kernel(cx, a + i * ST, b + i * FE, c + i * SE);
  } else if constexpr (OP == T_SSEED) {
    TL::template smooth_combine<true>(cx, a, b + i * SE, c + i * ST);
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
#ifdef POF_TUNE
  // tuning builds: the flag value is a time stamp (ns, low 31 bits) so that the host can read when each node completed
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  v = (unsigned)(t & 0x7fffffffu) | 0x80000000u;
#endif
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Kogge-Stone stage of the hybrid filter sweep (FlowArgs): where node j's window element lives and which flag says so
__device__ __forceinline__ int ks_nbits(long j) { return j == 0 ? 0 : 64 - __clzll(j); }
template <int EL>
__device__ __forceinline__ real* ks_elem(const FlowArgs& A, int l, long j) {
  return l == 0 ? A.agg + (A.off[A.ks_base] + j) * EL : A.ks + ((long)(l - 1) * A.ks_n + j) * EL;
}
__device__ __forceinline__ const unsigned* ks_flag(const FlowArgs& A, int l, long j) {
  if (l == 0) return (A.ks_base >= A.up_lo && A.ks_base <= A.up_hi) ? A.flag_up + A.off[A.ks_base] + j : nullptr;
  return A.ks_wait ? A.flag_ks + (long)(l - 1) * A.ks_n + j : nullptr;
}

template <int D, bool FILT, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_tree_flow(const FlowArgs A) {
  extern __shared__ __align__(16) real sm[];
  using TL = TreeLane<D>;
  constexpr int G = FILT ? TL::G2 : TL::GS;
  constexpr int CPW = 32 / G;
  constexpr int SMC = FILT ? TL::SM_COMBINE : TL::SM_COMBINE_S;
  constexpr int FE = 3 * D * D + 2 * D, SE = 2 * D * D + D, ST = D * D + D;
  constexpr int EL = FILT ? FE : SE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (A.stop && *A.stop != real(0)) return;  // the device-side IEKS loop has ended
  typename TL::Ctx cx;
  TL::template init<G>(cx, sm + (warp * CPW + lane / G) * SMC, FILT ? TL::NMAT : TL::NMAT_S);
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
    } else if (live && (kind == FlowArgs::DOWN || kind == FlowArgs::DOWN_E)) {
      dep0 = A.flag_dn + A.off[lev] + i;
      const long e = FILT ? 2 * i : 2 * i + 1;  // the child element the combine reads
      if (2 * i + 1 < A.sz[lev - 1] && lev - 1 >= A.up_lo && lev - 1 <= A.up_hi) dep1 = A.flag_up + A.off[lev - 1] + e;
    } else if (live && kind == FlowArgs::KS) {
      // step lev, in scan order r (filter: r = node, prefix scan; smoother: r counts from the LAST node, suffix scan):
      // node r = i + 2^(lev-1) combines the window ending at r' = i (as final as it is after step lev-1) with its own
      const long nb1 = A.ks_n - 1;
      const long ra = i, rj = i + (1L << (lev - 1));
      dep0 = ks_flag(A, min(lev - 1, ks_nbits(ra)), FILT ? ra : nb1 - ra);
      dep1 = ks_flag(A, lev - 1, FILT ? rj : nb1 - rj);
    } else if (live && kind == FlowArgs::KS_APPLY) {
      if constexpr (FILT) {
        dep0 = A.flag_dn + A.off[A.nlev - 1];
        if (i >= 1) dep1 = ks_flag(A, ks_nbits(i - 1), i - 1);
      } else {
        const long r = A.ks_n - 1 - i;  // node i needs the inclusive suffix of node i + 1 (scan position r - 1)
        if (r >= 1) dep0 = ks_flag(A, ks_nbits(r - 1), i + 1);
      }
    }
    while (true) {
      const bool ok = (!dep0 || ld_acquire(dep0) != 0u) && (!dep1 || ld_acquire(dep1) != 0u);
      if (__all_sync(0xffffffffu, ok)) break;
      if (A.poll_ns > 0) __nanosleep(A.poll_ns);
    }
    __syncwarp();
    if (live) {
      // ONE call site per combine form: the element-form and the state-form combine are ~4000 straight-line
      // instructions each; inlined once per item kind the kernel was 257 KB of code, and (measured with time-stamped
      // flags) every change of kind cost an extra ~8 us of cold instruction fetches on the critical path.
      const real* in1 = nullptr;   // element form: earlier (filter) / later (smoother) element; state form: the state
      const real* in2 = nullptr;
      real* outp = nullptr;
      int form = 0;                // 0: no combine, 1: element form, 2: state form
      const real* cp_src = nullptr;
      real* cp_dst = nullptr;
      int cp_n = 0;
      unsigned* rel0 = nullptr;
      unsigned* rel1 = nullptr;
      if (kind == FlowArgs::UP) {
        const real* lc = A.agg + (A.off[lev - 1] + 2 * i) * EL;
        real* pa = A.agg + (A.off[lev] + i) * EL;
        if (2 * i + 1 < A.sz[lev - 1]) {
          form = 1;
          in1 = FILT ? lc : lc + EL;
          in2 = FILT ? lc + EL : lc;
          outp = pa;
        } else {
          cp_src = lc, cp_dst = pa, cp_n = EL;
        }
        rel0 = A.flag_up + A.off[lev] + i;
      } else if (kind == FlowArgs::KS) {
        const long nb1 = A.ks_n - 1;
        const long ra = i, rj = i + (1L << (lev - 1));
        const long ja = FILT ? ra : nb1 - ra, j = FILT ? rj : nb1 - rj;
        form = 1;
        in1 = ks_elem<EL>(A, min(lev - 1, ks_nbits(ra)), ja);  // filter: the EARLIER window; smoother: the LATER one
        in2 = ks_elem<EL>(A, lev - 1, j);
        outp = ks_elem<EL>(A, lev, j);
        rel0 = A.flag_ks + (long)(lev - 1) * A.ks_n + j;
      } else if (kind == FlowArgs::KS_APPLY) {
        if constexpr (FILT) {
          const real* root = A.st + A.off[A.nlev - 1] * ST;
          real* out = A.st + (A.off[A.ks_base] + i) * ST;
          if (i == 0) {
            if (out != root) cp_src = root, cp_dst = out, cp_n = ST;
          } else {
            form = 2;
            in1 = root;
            in2 = ks_elem<EL>(A, ks_nbits(i - 1), i - 1);
            outp = out;
          }
        } else {
          // element-form suffix scan: a node's "everything later" aggregate is the inclusive suffix of the next node
          // (the identity element g = 0, E = I, D = 0 behind the last one)
          const long r = A.ks_n - 1 - i;
          real* out = A.sx + (A.off[A.ks_base] + i) * SE;
          if (r == 0) {
            for (int j = cx.r; j < SE; j += G) out[j] = (j >= D && j < D + D * D && (j - D) / D == (j - D) % D) ? 1.0 : 0.0;
          } else {
            cp_src = ks_elem<EL>(A, ks_nbits(r - 1), i + 1), cp_dst = out, cp_n = SE;
          }
        }
        rel0 = A.flag_dn + A.off[A.ks_base] + i;
      } else if (kind == FlowArgs::ROOT) {
        if (A.root_m) {
          real* r = A.st + A.off[A.nlev - 1] * ST;
          for (int j = cx.r; j < ST; j += G) r[j] = (j < D) ? __ldcg(A.root_m + j) : __ldcg(A.root_L + (j - D));
        } else if constexpr (!FILT) {
          // element-form down-sweep of the smoother: nothing lies later than the root -> the identity element
          // (g = 0, E = I, D = 0); combining it with any element returns that element exactly
          real* r = A.sx + A.off[A.nlev - 1] * SE;
          for (int j = cx.r; j < SE; j += G) r[j] = (j >= D && j < D + D * D && (j - D) / D == (j - D) % D) ? 1.0 : 0.0;
        }
        rel0 = A.flag_dn + A.off[A.nlev - 1];
      } else if (kind == FlowArgs::DOWN_E) {
        // exclusive-suffix ELEMENTS: X(right child) = X(parent); X(left child) = op(later = X(parent), right child)
        const real* px = A.sx + (A.off[lev] + i) * SE;
        const real* el = A.agg + A.off[lev - 1] * SE;
        real* cxs = A.sx + A.off[lev - 1] * SE;
        const bool two = 2 * i + 1 < A.sz[lev - 1];
        cp_src = px, cp_dst = cxs + (two ? 2 * i + 1 : 2 * i) * SE, cp_n = SE;
        if (two) {
          form = 1;
          in1 = px;
          in2 = el + (2 * i + 1) * SE;
          outp = cxs + 2 * i * SE;
          rel1 = A.flag_dn + A.off[lev - 1] + 2 * i + 1;
        }
        rel0 = A.flag_dn + A.off[lev - 1] + 2 * i;
      } else {  // DOWN, state form
        const real* p = A.st + (A.off[lev] + i) * ST;
        const real* el = A.agg + A.off[lev - 1] * EL;
        real* cs = A.st + A.off[lev - 1] * ST;
        const bool two = 2 * i + 1 < A.sz[lev - 1];
        // filter: the left child starts where its parent starts; smoother: the right child ends where its parent ends
        cp_src = p, cp_dst = cs + ((FILT || !two) ? 2 * i : 2 * i + 1) * ST, cp_n = ST;
        if (two) {
          form = 2;
          in1 = p;
          in2 = el + (FILT ? 2 * i : 2 * i + 1) * EL;
          outp = cs + (FILT ? 2 * i + 1 : 2 * i) * ST;
          rel1 = A.flag_dn + A.off[lev - 1] + 2 * i + 1;
        }
        rel0 = A.flag_dn + A.off[lev - 1] + 2 * i;
      }
      if (cp_n) group_copy<D, G>(cx.r, cp_dst, cp_src, cp_n);
      if (form == 1) {
        if constexpr (FILT) TL::template filter_combine<false>(cx, in1, in2, outp);
        else TL::template smooth_combine<false>(cx, in1, in2, outp);
      } else if (form == 2) {
        if constexpr (FILT) TL::template filter_combine<true>(cx, in1, in2, outp);
        else TL::template smooth_combine<true>(cx, in1, in2, outp);
      }
      __threadfence();
      cx.sync();
      if (cx.r == 0) {
        if (rel0) st_release(rel0, 1u);
        if (rel1) st_release(rel1, 1u);
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Rank-carry exchanges of the time-sharded pass (pof/sharded.py): what every rank does with the all-gathered carries,
// ONE launch of one lane group per exchange (register-resident state-form combines, ~3 us each), including the
// scalar bookkeeping that used to be ~25 small torch kernels per iteration.
//   filter  : state_in = x0 (+) carry_0 (+) ... (+) carry_{rank-1}           (filter.py:117-142 in state form)
//   smoother: sums of the ranks' partial statistics in rank order (bitwise identical on every rank) -> nll, sigma^2,
//             calibration scale; seed = terminal state (last rank's end state) combined with the later ranks'
//             smoothing carries W-1 .. rank+1                                 (smoother.py:53-63 in state form)
template <int D, bool FILT>
__global__ void __launch_bounds__(32) k_exchange(ExchangeArgs A) {
  extern __shared__ __align__(16) real sm[];
  using TL = TreeLane<D>;
  constexpr int G = FILT ? TL::G2 : TL::GS;
  constexpr int FE = 3 * D * D + 2 * D, SE = 2 * D * D + D, ST = D * D + D;
  const int lane = threadIdx.x & 31;
  if (A.p2p) {
    static_assert(sizeof(real) == 8 || sizeof(real) == 4, "");
    // filter: the carries of the earlier ranks; smoother: every rank's payload (the partial sums of all are needed)
    A.gathered = FILT ? p2p_exchange(A, 0, A.rank) : p2p_exchange(A, 0, A.world);
  }
  if (lane >= G) return;
  typename TL::Ctx cx;
  TL::template init<G>(cx, sm, FILT ? TL::NMAT : TL::NMAT_S);
  if constexpr (FILT) {
    // pack x0 into the scratch/state ping-pong so that the last combine writes state_out
    const int count = A.rank;
    real* bufs[2] = {A.state_out, A.scratch};
    int cur = count & 1;  // after `count` flips the result sits in bufs[0]
    for (int j = cx.r; j < ST; j += G) bufs[cur][j] = (j < D) ? __ldcg(A.x0_mean + j) : __ldcg(A.x0_chol + (j - D));
    __threadfence();
    cx.sync();
    for (int i = 0; i < count; ++i) {
      TL::template filter_combine<true>(cx, bufs[cur], A.gathered + (long)i * A.stride, bufs[cur ^ 1]);
      __threadfence();
      cx.sync();
      cur ^= 1;
    }
  } else {
    // gathered payload of rank r: [smoothing carry SE | filtered end state ST | nll, s1, s2 partial sums]
    const int W = A.world;
    if (cx.r == 0) {
      real s0 = 0.0, s1 = 0.0, s2 = 0.0;
      for (int r = 0; r < W; ++r) {
        const real* p = A.gathered + (long)r * A.stride + SE + ST;
        s0 += __ldcg(p);
        s1 += __ldcg(p + 1);
        s2 += __ldcg(p + 2);
      }
      const real ssq = s1 / A.n_obs / A.d_obs;
      const real cs = A.calibrate ? sqrt(ssq) : 1.0;
      *A.cscale = cs;
      if (A.scalars) {
        A.scalars[0] = s0;                       // POF_S_NLL
        A.scalars[2] = ssq;                      // POF_S_SSQ
        A.scalars[3] = s2 / A.n_obs / A.d_obs;   // POF_S_SSQ_PROPER
        A.scalars[5] = cs;                       // POF_S_CSCALE
      }
    }
    const int count = W - 1 - A.rank;
    real* bufs[2] = {A.state_out, A.scratch};
    int cur = count & 1;
    const real* term = A.gathered + (long)(W - 1) * A.stride + SE;
    for (int j = cx.r; j < ST; j += G) bufs[cur][j] = __ldcg(term + j);
    __threadfence();
    cx.sync();
    for (int i = 0; i < count; ++i) {  // later carries first: W-1, W-2, .., rank+1
      const real* el = A.gathered + (long)(W - 1 - i) * A.stride;
      TL::template smooth_combine<true>(cx, bufs[cur], el, bufs[cur ^ 1]);
      __threadfence();
      cx.sync();
      cur ^= 1;
    }
  }
}

template <int D>
struct TreeLaunchers {
  using TL = TreeLane<D>;
  template <int OP>
  static cudaError_t run(cudaStream_t s, const real* a, long na, const real* b, real* c, long nb) {
    constexpr bool FILT = (OP == T_FUP || OP == T_FDOWN || OP == T_FCOMB || OP == T_CHUNKK);
    constexpr int G = FILT ? TL::G2 : TL::GS;
    constexpr int CPW = 32 / G;
    // the ops that run on the side stream (next to the resident CTAs of the filter scan) use two-warp CTAs
    constexpr int WARPS = (OP == T_CHUNKK || OP == T_SUP) ? 2 : TL_WARPS;
    constexpr int smem = WARPS * CPW * (FILT ? TL::SM_COMBINE : TL::SM_COMBINE_S) * (int)sizeof(real);
    if (nb <= 0) return cudaSuccess;
    if (cudaError_t e = ensure_smem(k_tree<D, OP, WARPS>, smem, WARPS == 2)) return e;
    const long per_block = (long)WARPS * CPW;
    k_tree<D, OP, WARPS><<<(unsigned)((nb + per_block - 1) / per_block), WARPS * 32, smem, s>>>(a, na, b, c, nb);
    return cudaGetLastError();
  }
  // dataflow whole-sweep launch: persistent grid (enough warps for the widest level, at most what is resident)
  template <bool FILT>
  static cudaError_t flow(cudaStream_t s, const FlowArgs& A) {
    constexpr int G = FILT ? TL::G2 : TL::GS;
    constexpr int CPW = 32 / G;
    constexpr int WARPS = FILT ? TL_WARPS : 2;  // smoother sweeps run next to the filter scan's resident CTAs
    constexpr int smem = WARPS * CPW * (FILT ? TL::SM_COMBINE : TL::SM_COMBINE_S) * (int)sizeof(real);
    auto kern = k_tree_flow<D, FILT, WARPS>;
    if (cudaError_t e = ensure_smem(kern, smem, !FILT)) return e;
    static int max_grid[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
    if (max_grid[dev] == 0) {
      int per_sm = 0, sms = 0;
      if (cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, WARPS * 32, smem)) return e;
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
    long blocks = (widest + (long)WARPS * CPW - 1) / ((long)WARPS * CPW);
    if (blocks < 1) blocks = 1;
    if (blocks > max_grid[dev]) blocks = max_grid[dev];
    k_tree_flow<D, FILT, WARPS><<<(unsigned)blocks, WARPS * 32, smem, s>>>(a);
    return cudaGetLastError();
  }
  template <bool FILT>
  static cudaError_t exchange(cudaStream_t s, const ExchangeArgs& A) {
    constexpr int smem = (FILT ? TL::SM_COMBINE : TL::SM_COMBINE_S) * (int)sizeof(real);
    if (cudaError_t e = ensure_smem(k_exchange<D, FILT>, smem)) return e;
    k_exchange<D, FILT><<<1, 32, smem, s>>>(A);
    return cudaGetLastError();
  }
  static const TreeLaunch* get() {
    static const TreeLaunch t = {&run<T_FUP>, &run<T_FDOWN>, &run<T_SUP>, &run<T_SDOWN>, &run<T_FCOMB>, &run<T_SCOMB>,
                                 &run<T_CHUNKK>, &run<T_SSEED>, &flow<true>, &flow<false>, &exchange<true>,
                                 &exchange<false>};
    return &t;
  }
};

}  // namespace POF_NS
