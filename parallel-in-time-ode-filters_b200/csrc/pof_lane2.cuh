// Leaf level, performance path v2: G lanes of a warp cooperate on ONE time-chunk and every lane owns R = 2 ROWS of each
// D-row matrix (row = lane + s*G, s = 0,1), G = pow2ceil(ceil(D/2)): D = 8 -> 4 lanes per chunk, 8 chunks per warp.
//
// Why two rows per lane (measured on B200 with the one-row kernels of pof_lane.cuh, profiles/r01_*): every
// Householder pivot has to deliver its K+1 pivot-row entries to every lane, 2 crossbar wavefronts per real per warp
// (shared memory or shuffle alike), while a lane with one row does only 2(K+1) FMAs with them -- the kernels sat at
// ~35 % FP64-pipe utilisation with the shared-memory/shuffle crossbar 55-65 % busy and 8 warps per SM (255 registers)
// leaving dependency ("wait") stalls exposed.  With two rows per lane the same broadcast feeds twice the FMAs, the two
// row updates are independent instruction streams (ILP), the redundant reflector set-up is shared by twice as many
// chunks per warp, and rows that lie above the pivot for a whole slot are skipped at compile time.
//
// Further specialisations relative to pof_lane.cuh (same math, see pof_leaf.cuh / pof_pipeline.cuh for citations):
//   * KC: the posterior factor of a noiseless update has d zero columns -> the prediction QR only carries D-d columns
//     (all steps but a chunk's first in the scan; the fold's first step has no factor at all).
//   * compact linearisation: H = E1 - J_f E0 has d+1 non-zeros per row -> H T, H A, H m cost (d+1) terms, and H is
//     never held in registers (dense H, the S3 seam, still supported).
#pragma once
#include <cuda_runtime.h>

#include "pof_real.cuh"
#include "pof_small.cuh"

namespace POF_NS {
// binomial coefficients of the Pascal blocks in the scalar type of this build
__host__ __device__ constexpr real binomr(int n, int k) { return (real)pof::binom(n, k); }

template <int d, int q>
struct Lane2 {
  static constexpr int Q1 = q + 1;
  static constexpr int D = d * Q1;
  static constexpr int R = 2;
  static constexpr int pow2c(int x) { return x <= 1 ? 1 : (x <= 2 ? 2 : (x <= 4 ? 4 : (x <= 8 ? 8 : 16))); }
  static constexpr int G = pow2c((D + R - 1) / R);
  static constexpr int GPW = 32 / G;
  static constexpr int NE = D + 2 * D * D;
  static constexpr int KP = D - d;  // columns of a posterior factor
  // The information factor Z of a chunk's filtering element only ACCUMULATES: Z Z^T += G_k^T G_k per step, nothing
  // else in the fold reads it.  So the d new columns of MZ consecutive steps are collected first and absorbed with
  // ONE triangularisation tria([Z | G_k^T ... G_{k+MZ-1}^T]) (D pivots with MZ*d extra columns) instead of MZ
  // triangularisations with d extra columns each: the reflector set-up (norm, rsqrt, rcp: as expensive as the row
  // update when there are only d columns) and the dependent pivot chain shrink by the factor MZ.
#ifndef POF_FOLD_MZ
#define POF_FOLD_MZ 4
#endif
  static constexpr int MZ = (POF_FOLD_MZ * d <= D) ? POF_FOLD_MZ : (D / d);
  static constexpr real LOG_2PI = 1.8378770664093454835606594728112;
  // per-group shared memory (doubles): two row-exchange matrices, two gather vectors, this group's rows of QL
  static constexpr int LDM = D + 1;
  static constexpr int VEC = ((D + 1) / 2) * 2;
  static constexpr int NJ = d * d + d;                 // compact linearisation of a step: [J_f | c]
  static constexpr int NJP = ((NJ + 1) / 2) * 2;
  static constexpr int RAW = 2 * D * LDM + 2 * VEC + R * G * D + 2 * NJP;  // ... + 2 staging slots for [J_f | c]
  // 64-bit shared accesses are served per half-warp (16 lanes = 16/G groups): the group stride must spread those
  // groups over the 16 eight-byte banks, i.e. SM_GROUP == G (mod 16) (with odd LDM the rows of a group then fall into
  // distinct banks as well).  ncu before: 2x excess wavefronts on every STS and on the column reads (stride == 2).
  static constexpr int SM_GROUP = RAW + (((G % 16) - (RAW % 16)) + 16) % 16;

  // ---- smoother, OPT-IN (flag POF_F_SMOOTH_TMA): bulk-copy (TMA engine) staging of the next step's backward kernel.
  // One step's kernel (g | E | noise factor: NE contiguous values) is copied global -> shared by ONE cp.async.bulk per
  // chunk while the current step computes, its completion is signalled on an mbarrier that belongs to the chunk's lane
  // group; the step then reads its rows with LDS (168 instead of 255 registers).  Built to remove the 14 LDG.128 per
  // lane and step (ncu: long_scoreboard 9 % + mio_throttle 8 % of the smoother's stall samples).  MEASURED on B200:
  // correct, but 30x SLOWER (13.9 ms instead of 0.455 ms at N = 2^20, with or without the proxy fence, blocking or
  // non-blocking wait): 64 independent 1 KB streams per SM = ~1e6 bulk copies per pass complete at only ~0.5 copies
  // per microsecond and SM -- the copy engine is made for few large tiles, not for many small ones.  Hence off by
  // default; kept because the negative result is the evidence.  Needs 16-byte aligned sizes; block = NE values + mbarrier.
  static constexpr int STG = ((NE * (int)sizeof(real) + 8 + 15) / 16) * 16 / (int)sizeof(real);
  static constexpr bool STG_OK = (NE * sizeof(real)) % 16 == 0 && (STG * sizeof(real)) % 16 == 0;

  struct Lin {
    const real* __restrict__ H;
    const real* __restrict__ c;
    const real* __restrict__ Jc;
    real s0, s1;
  };

  struct Ctx {
    int l;             // lane within the group
    int row[R];        // rows owned (may be >= D: idle slot)
    int rc[R];         // clamped row index for addressing
    int rb[R], blk0[R];
    unsigned mask;
    real* mat;       // 2 exchange matrices
    real* vec;       // 2 gather vectors
    real* tq;        // this lane's rows of QL, lane-minor: tq[(s*D + j)*G] (conflict-free across the group)
    real* lbuf;      // 2 x NJP doubles: compact linearisations staged by cp.async (global -> shared, no registers)
    real* stg;       // smoother with bulk-copy staging: this group's staging block (NE values, then the mbarrier), or null
    unsigned stg_phase;
    int vflip;
    real cf[R][Q1];  // Pascal coefficients of the owned rows of F
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
  };

  static __device__ __forceinline__ void init_ctx(Ctx& c, real* sm_group, const real* qL) {
    const int lane = threadIdx.x & 31;
    c.l = lane % G;
    c.mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
    c.mat = sm_group;
    c.vec = sm_group + 2 * D * LDM;
    c.tq = c.vec + 2 * VEC + c.l;
    c.lbuf = c.vec + 2 * VEC + R * G * D;
    c.stg = nullptr;
    c.stg_phase = 0;
    c.vflip = 0;
#pragma unroll
    for (int s = 0; s < R; ++s) {
      c.row[s] = c.l + s * G;
      c.rc[s] = (c.row[s] < D) ? c.row[s] : 0;
      c.rb[s] = c.rc[s] % Q1;
      c.blk0[s] = c.rc[s] - c.rb[s];
#pragma unroll
      for (int i = 0; i < Q1; ++i) {
        real v = 0.0;
#pragma unroll
        for (int b = 0; b < Q1; ++b)
          if (c.rb[s] == b && i >= b) v = binomr(q - b, i - b);
        c.cf[s][i] = v;
      }
#pragma unroll
      for (int j = 0; j < D; ++j) {
        real v = 0.0;
#pragma unroll
        for (int b = 0; b < Q1; ++b)
          if (c.rb[s] == b && (j / Q1) * Q1 == c.blk0[s] && (j % Q1) <= b) v = qL[b * Q1 + (j % Q1)];
        c.tq[(s * D + j) * G] = (c.row[s] < D) ? v : 0.0;
      }
    }
    c.sync();
  }

  // ---------------------------------------------------------------------------------------------- small helpers
  static __device__ __forceinline__ real fast_rcp(real x) { return POF_NS::fast_rcp(x); }
  static __device__ __forceinline__ real fast_rsqrt(real x) { return POF_NS::fast_rsqrt(x); }
  // sum_j a[j]*b[j] with several independent accumulators: the leaf recursions are bound by dependent-issue latency
  // ("wait" stalls with 2 warps per scheduler), so every long FMA chain is split and tree-added
  template <int n>
  static __device__ __forceinline__ real dotn(const real* a, const real* b) {
    if constexpr (n <= 0) {
      return 0.0;
    } else if constexpr (n < 4) {
      real s = a[0] * b[0];
#pragma unroll
      for (int j = 1; j < n; ++j) s = fma(a[j], b[j], s);
      return s;
    } else if constexpr (n < 10) {
      real s0 = a[0] * b[0], s1 = a[1] * b[1];
#pragma unroll
      for (int j = 2; j + 1 < n; j += 2) {
        s0 = fma(a[j], b[j], s0);
        s1 = fma(a[j + 1], b[j + 1], s1);
      }
      if constexpr (n % 2) s0 = fma(a[n - 1], b[n - 1], s0);
      return s0 + s1;
    } else {
      real s0 = a[0] * b[0], s1 = a[1] * b[1], s2 = a[2] * b[2], s3 = a[3] * b[3];
#pragma unroll
      for (int j = 4; j + 3 < n; j += 4) {
        s0 = fma(a[j], b[j], s0);
        s1 = fma(a[j + 1], b[j + 1], s1);
        s2 = fma(a[j + 2], b[j + 2], s2);
        s3 = fma(a[j + 3], b[j + 3], s3);
      }
      if constexpr (n % 4 >= 1) s0 = fma(a[n - n % 4], b[n - n % 4], s0);
      if constexpr (n % 4 >= 2) s1 = fma(a[n - n % 4 + 1], b[n - n % 4 + 1], s1);
      if constexpr (n % 4 >= 3) s2 = fma(a[n - n % 4 + 2], b[n - n % 4 + 2], s2);
      return (s0 + s1) + (s2 + s3);
    }
  }
  struct HH {
    real s, tp, beta;
  };
  // Householder for the row (alpha, x[0..n)):  H = I - tp v v^T, v = (s, x), H (alpha, x)^T = (beta, 0)
  template <int n>
  static __device__ __forceinline__ HH house(real alpha, const real* x) {
    const real sigma = dotn<n>(x, x);
    const real nrm2 = fma(alpha, alpha, sigma);
    // rsqrt.approx.ftz flushes subnormal inputs to zero (-> inf -> NaN in the Newton step): a row whose squared norm
    // is below 2^-1000 is treated as already reduced (identity reflector), as a zero tail is
    const bool nz = sigma > real(0) && nrm2 > tiny_norm2();
    const real rn = fast_rsqrt(nrm2);
    const real nrm = nrm2 * rn;
    const real beta = (alpha >= 0.0) ? -nrm : nrm;
    const real s = alpha - beta;
    HH h;
    h.beta = nz ? beta : alpha;
    h.s = nz ? s : 0.0;
    h.tp = nz ? rn * fast_rcp(fabs(s)) : 0.0;
    return h;
  }
  static __device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  }
  static __device__ __forceinline__ real bshfl(const Ctx& c, real x, int src) {
    return __shfl_sync(c.mask, x, ((threadIdx.x & 31) / G) * G + src);
  }
  // out[row] = x[s] of the lane owning `row`
  static __device__ __forceinline__ void gather(Ctx& c, const real (&x)[R], real (&out)[D]) {
    real* v = c.vec + c.vflip * VEC;
    c.vflip ^= 1;
#pragma unroll
    for (int s = 0; s < R; ++s)
      if (c.row[s] < D) v[c.row[s]] = x[s];
    c.sync();
#pragma unroll
    for (int j = 0; j < D; ++j) out[j] = v[j];
  }
  // publish the owned rows (columns [J0, D)) into exchange matrix `which`
  template <int J0>
  static __device__ __forceinline__ const real* publish(Ctx& c, int which, const real (&x)[R][D]) {
    real* M = c.mat + which * D * LDM;
    c.sync();
#pragma unroll
    for (int s = 0; s < R; ++s)
      if (c.row[s] < D) {
#pragma unroll
        for (int j = J0; j < D; ++j) M[c.row[s] * LDM + j] = x[s][j];
      }
    c.sync();
    return M;
  }
  // y[s][j] = (F X)[row_s][j], j in [J0, D), from the published rows of X
  template <int J0>
  static __device__ __forceinline__ void mulF_rows(const Ctx& c, const real* M, real (&y)[R][D]) {
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int j = J0; j < D; ++j) y[s][j] = 0.0;
#pragma unroll
      for (int i = 0; i < Q1; ++i) {
        const real* rowp = M + (c.blk0[s] + i) * LDM;
#pragma unroll
        for (int j = J0; j < D; ++j) y[s][j] = fma(c.cf[s][i], rowp[j], y[s][j]);
      }
    }
  }
  static __device__ __forceinline__ void mulF_vec(real (&m)[D]) {
#pragma unroll
    for (int b = 0; b < d; ++b) {
#pragma unroll
      for (int i = 0; i < Q1; ++i) {
#pragma unroll
        for (int j = i + 1; j < Q1; ++j) m[b * Q1 + i] = fma(binomr(q - i, j - i), m[b * Q1 + j], m[b * Q1 + i]);
      }
    }
  }
  static __device__ __forceinline__ real pick(const real (&x)[D], int idx) {
    real v = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i)
      if (i == idx) v = x[i];
    return v;
  }
  static __device__ __forceinline__ void load_tq(const Ctx& c, real (&t)[R][D]) {
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int j = 0; j < D; ++j) t[s][j] = c.tq[(s * D + j) * G];
    }
  }

  // ---------------------------------------------------------------------------------------------- triangularisations
  // Triangular-pentagonal right-QR over the columns [J0, D) of C.  t: rows of T (lower triangular), c: rows of C.
  // Passenger rows pt (D entries), pc (columns [J0, D)) see the same reflections.
  // (the per-slot update is a template on the slot so that slots lying entirely above the pivot vanish at compile
  // time)
  template <int I, int J0, bool PASS>
  static __device__ __forceinline__ void tp_step2(Ctx& cx, real (&t)[R][D], real (&c)[R][D], real (*pt)[D],
                                                  real (*pc)[D]) {
    if constexpr (I < D) {
      constexpr int K = D - J0;
      constexpr int so = I / G;
      const int lo = I % G;
      real piv[K + 1];
      piv[0] = bshfl(cx, t[so][I], lo);
#pragma unroll
      for (int j = 0; j < K; ++j) piv[1 + j] = bshfl(cx, c[so][J0 + j], lo);
      const HH h = house<K>(piv[0], piv + 1);
      row_update<0, I, J0, K>(cx, h, piv, t, c);
      row_update<1, I, J0, K>(cx, h, piv, t, c);
      if constexpr (PASS) {
#pragma unroll
        for (int s = 0; s < R; ++s) {
          real u = fma(h.s, pt[s][I], dotn<K>(&pc[s][J0], piv + 1));
          u *= h.tp;
          pt[s][I] = fma(-u, h.s, pt[s][I]);
#pragma unroll
          for (int j = 0; j < K; ++j) pc[s][J0 + j] = fma(-u, piv[1 + j], pc[s][J0 + j]);
        }
      }
      tp_step2<I + 1, J0, PASS>(cx, t, c, pt, pc);
    }
  }
  template <int S, int I, int J0, int K>
  static __device__ __forceinline__ void row_update(const Ctx& cx, const HH& h, const real* piv, real (&t)[R][D],
                                                    real (&c)[R][D]) {
    if constexpr (S * G + G - 1 >= I) {  // some row of this slot is at or below the pivot
      real w = fma(h.s, t[S][I], dotn<K>(&c[S][J0], piv + 1));  // the dot does not wait for the reflector
      w = (cx.row[S] >= I) ? w * h.tp : 0.0;
      t[S][I] = (cx.row[S] == I) ? h.beta : fma(-w, h.s, t[S][I]);
#pragma unroll
      for (int j = 0; j < K; ++j) c[S][J0 + j] = fma(-w, piv[1 + j], c[S][J0 + j]);
    }
  }
  template <int J0, bool PASS>
  static __device__ __forceinline__ void tpqrt(Ctx& cx, real (&t)[R][D], real (&c)[R][D], real (*pt)[D],
                                               real (*pc)[D]) {
    tp_step2<0, J0, PASS>(cx, t, c, pt, pc);
  }

  // plain right-Householder lower-triangularisation of the D x (D - J0) matrix held in columns [J0, D) of x; the
  // result is written as a lower-triangular D x D factor into columns [0, D - J0) (pivot I uses column J0 + I)
  template <int S, int I, int J0>
  static __device__ __forceinline__ void tria_update(const Ctx& cx, const HH& h, const real* piv,
                                                     real (&x)[R][D]) {
    if constexpr (S * G + G - 1 >= I) {
      constexpr int n = D - J0 - I;
      real w = fma(h.s, x[S][J0 + I], dotn<n - 1>(&x[S][J0 + I + 1], piv + 1));
      w = (cx.row[S] >= I) ? w * h.tp : 0.0;
      x[S][J0 + I] = (cx.row[S] == I) ? h.beta : fma(-w, h.s, x[S][J0 + I]);
#pragma unroll
      for (int j = 1; j < n; ++j) x[S][J0 + I + j] = fma(-w, piv[j], x[S][J0 + I + j]);
    }
  }
  template <int I, int J0>
  static __device__ __forceinline__ void tria_step(Ctx& cx, real (&x)[R][D]) {
    if constexpr (I + 1 < D - J0) {
      constexpr int n = D - J0 - I;
      constexpr int so = I / G;
      const int lo = I % G;
      real piv[n];
#pragma unroll
      for (int j = 0; j < n; ++j) piv[j] = bshfl(cx, x[so][J0 + I + j], lo);
      const HH h = house<n - 1>(piv[0], piv + 1);
      tria_update<0, I, J0>(cx, h, piv, x);
      tria_update<1, I, J0>(cx, h, piv, x);
      tria_step<I + 1, J0>(cx, x);
    }
  }
  // x (columns [J0, D)) -> lower-triangular factor in out (columns [0, D)), zero above the diagonal / beyond D-J0
  template <int J0>
  static __device__ __forceinline__ void tria_rows(Ctx& cx, real (&x)[R][D], real (&out)[R][D]) {
    tria_step<0, J0>(cx, x);
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int j = 0; j < D; ++j) out[s][j] = (j < D - J0 && j <= cx.row[s]) ? x[s][J0 + j] : 0.0;
    }
  }

  // Right-Householder lower-triangularisation of the D x (2D - JS) matrix [x | y(:, JS:D)]; the D x D lower-triangular
  // result is left in x.  (Smoother: x = E L_{k+1}, y = the UNtriangularised Phi22~ block of the step's joint QR --
  // only y y^T matters, so the filter scan does not triangularise it.)
  template <int S, int I, int JS>
  static __device__ __forceinline__ void tria2_update(const Ctx& cx, const HH& h, const real* piv, real (&x)[R][D],
                                                      real (&y)[R][D]) {
    if constexpr (S * G + G - 1 >= I) {
      constexpr int n1 = D - I, n2 = D - JS;
      real w = fma(h.s, x[S][I], dotn<n1 - 1>(&x[S][I + 1], piv + 1) + dotn<n2>(&y[S][JS], piv + n1));
      w = (cx.row[S] >= I) ? w * h.tp : 0.0;
      x[S][I] = (cx.row[S] == I) ? h.beta : fma(-w, h.s, x[S][I]);
#pragma unroll
      for (int j = 1; j < n1; ++j) x[S][I + j] = fma(-w, piv[j], x[S][I + j]);
#pragma unroll
      for (int j = 0; j < n2; ++j) y[S][JS + j] = fma(-w, piv[n1 + j], y[S][JS + j]);
    }
  }
  template <int I, int JS>
  static __device__ __forceinline__ void tria2_step(Ctx& cx, real (&x)[R][D], real (&y)[R][D]) {
    if constexpr (I < D) {
      constexpr int n1 = D - I, n2 = D - JS;
      constexpr int so = I / G;
      const int lo = I % G;
      real piv[n1 + n2];
#pragma unroll
      for (int j = 0; j < n1; ++j) piv[j] = bshfl(cx, x[so][I + j], lo);
#pragma unroll
      for (int j = 0; j < n2; ++j) piv[n1 + j] = bshfl(cx, y[so][JS + j], lo);
      const HH h = house<n1 + n2 - 1>(piv[0], piv + 1);
      tria2_update<0, I, JS>(cx, h, piv, x, y);
      tria2_update<1, I, JS>(cx, h, piv, x, y);
      tria2_step<I + 1, JS>(cx, x, y);
    }
  }

  // ---------------------------------------------------------------------------------------------- linearisation access
  struct LinK {  // one step's linearisation, compact or dense
    real J[d][d], c[d];
    const real* Hd;  // dense H of this step, or null
  };
  static __device__ __forceinline__ void load_lin(const Lin& L, long k, LinK& o) {
    if (L.Jc) {
      const real* p = L.Jc + k * (d * d + d);
#pragma unroll
      for (int a = 0; a < d; ++a) {
        o.c[a] = __ldg(p + d * d + a);
#pragma unroll
        for (int b = 0; b < d; ++b) o.J[a][b] = __ldg(p + a * d + b);
      }
      o.Hd = nullptr;
    } else {
#pragma unroll
      for (int a = 0; a < d; ++a) {
        o.c[a] = __ldg(L.c + k * d + a);
#pragma unroll
        for (int b = 0; b < d; ++b) o.J[a][b] = 0.0;
      }
      o.Hd = L.H + k * d * D;
    }
  }
  // bring the linearisation of step k into L2 (the load itself sits next to its use: held in registers across the
  // prediction QR it was spilled right after the load, which exposed the full DRAM latency -- ncu: STL on long_sb)
  static __device__ __forceinline__ void prefetch_lin(const Lin& L, long k) {
    if (L.Jc) {
      const real* p = L.Jc + k * (d * d + d);
      prefetch_l2(p);
      prefetch_l2(p + (d * d + d) - 1);
    } else {
      prefetch_l2(L.c + k * d);
      prefetch_l2(L.H + k * d * D);
      prefetch_l2(L.H + k * d * D + d * D - 1);
    }
  }
  // Asynchronous staging of step k's compact linearisation into slot (k & 1) of the group's shared memory
  // (cp.async: global -> shared without passing through registers; held in registers one step ahead it cost ~60
  // registers of pressure and spills, loaded at its use it exposed the DRAM latency once per step).
  static constexpr int CPD = (G >= 2 && NJ % 2 == 0) ? 2 : 1;  // doubles per copy (16-byte copies need alignment)
  static __device__ __forceinline__ void stage_lin(const Ctx& cx, const Lin& L, long k, bool valid) {
    if (L.Jc && valid) {
      const real* src = L.Jc + k * NJ;
      real* dst = cx.lbuf + (k & 1) * NJP;
#pragma unroll
      for (int i0 = 0; i0 < NJ / CPD; i0 += G) {
        const int i = i0 + cx.l;
        if (i < NJ / CPD) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + i * CPD);
          constexpr int BYTES = CPD * (int)sizeof(real);
          if constexpr (BYTES == 16)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + i * CPD) : "memory");
          else if constexpr (BYTES == 8)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(src + i * CPD) : "memory");
          else
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(src + i * CPD) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // step k's linearisation: from the staging slot (waits for all but the most recent cp.async group), or dense
  static __device__ __forceinline__ void fetch_lin(const Ctx& cx, const Lin& L, long k, LinK& o) {
    if (L.Jc) {
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      cx.sync();
      const real* p = cx.lbuf + (k & 1) * NJP;
#pragma unroll
      for (int a = 0; a < d; ++a) {
        o.c[a] = p[d * d + a];
#pragma unroll
        for (int b = 0; b < d; ++b) o.J[a][b] = p[a * d + b];
      }
      o.Hd = nullptr;
    } else {
      load_lin(L, k, o);
    }
  }
  // (H X)[a][col] for a column `col` of a published D-row matrix M
  static __device__ __forceinline__ void H_times_col(const Lin& L, const LinK& lk, const real* M, int col,
                                                     real (&out)[d]) {
    if (lk.Hd == nullptr) {
#pragma unroll
      for (int a = 0; a < d; ++a) {
        real s = L.s1 * M[(a * Q1 + 1) * LDM + col];
#pragma unroll
        for (int b = 0; b < d; ++b) s = fma(-L.s0 * lk.J[a][b], M[(b * Q1) * LDM + col], s);
        out[a] = s;
      }
    } else {
#pragma unroll
      for (int a = 0; a < d; ++a) out[a] = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const real v = M[i * LDM + col];
#pragma unroll
        for (int a = 0; a < d; ++a) out[a] = fma(__ldg(lk.Hd + a * D + i), v, out[a]);
      }
    }
  }
  // H v + c for a replicated vector v
  static __device__ __forceinline__ void H_times_vec(const Lin& L, const LinK& lk, const real (&v)[D],
                                                     real (&out)[d]) {
    if (lk.Hd == nullptr) {
#pragma unroll
      for (int a = 0; a < d; ++a) {
        real s = fma(L.s1, v[a * Q1 + 1], lk.c[a]);
#pragma unroll
        for (int b = 0; b < d; ++b) s = fma(-L.s0 * lk.J[a][b], v[b * Q1], s);
        out[a] = s;
      }
    } else {
#pragma unroll
      for (int a = 0; a < d; ++a) {
        real s = lk.c[a];
#pragma unroll
        for (int i = 0; i < D; ++i) s = fma(__ldg(lk.Hd + a * D + i), v[i], s);
        out[a] = s;
      }
    }
  }

  // ---------------------------------------------------------------------------------------------- measurement update
  template <int A>
  static __device__ __forceinline__ void update_pivot(real (&t)[R][D], real (&W)[d][D]) {
    if constexpr (A < d) {
      const HH h = house<D - A - 1>(W[A][A], &W[A][A + 1]);
#pragma unroll
      for (int s = 0; s < R; ++s) {
        real w = fma(h.s, t[s][A], dotn<D - A - 1>(&t[s][A + 1], &W[A][A + 1]));
        w *= h.tp;
        t[s][A] = fma(-w, h.s, t[s][A]);
#pragma unroll
        for (int j = A + 1; j < D; ++j) t[s][j] = fma(-w, W[A][j], t[s][j]);
      }
#pragma unroll
      for (int a2 = A + 1; a2 < d; ++a2) {
        real u = fma(h.s, W[a2][A], dotn<D - A - 1>(&W[a2][A + 1], &W[A][A + 1]));
        u *= h.tp;
        W[a2][A] = fma(-u, h.s, W[a2][A]);
#pragma unroll
        for (int j = A + 1; j < D; ++j) W[a2][j] = fma(-u, W[A][j], W[a2][j]);
      }
      W[A][A] = h.beta;
      update_pivot<A + 1>(t, W);
    }
  }
  // In: t = rows of the predicted factor T (lower triangular, zeros above the diagonal).  Out: SLinv-ready SL, and
  // t = [Kbar | posterior factor] rows.  M: T must have been published into exchange matrix 0 by the caller.
  static __device__ __forceinline__ void update(Ctx& cx, const Lin& L, const LinK& lk, const real* MT,
                                                real (&t)[R][D], real (&SL)[d][d]) {
    real wc[d][R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      real o[d];
      H_times_col(L, lk, MT, cx.rc[s], o);
#pragma unroll
      for (int a = 0; a < d; ++a) wc[a][s] = o[a];
    }
    real W[d][D];
#pragma unroll
    for (int a = 0; a < d; ++a) gather(cx, wc[a], W[a]);
    update_pivot<0>(t, W);
#pragma unroll
    for (int a = 0; a < d; ++a) {
#pragma unroll
      for (int e = 0; e < d; ++e) SL[a][e] = (e <= a) ? W[a][e] : 0.0;
    }
  }
  static __device__ __forceinline__ void solveSL(const real (&SL)[d][d], const real (&y)[d], real (&z)[d]) {
#pragma unroll
    for (int a = 0; a < d; ++a) {
      real s = y[a];
#pragma unroll
      for (int j = 0; j < a; ++j) s = fma(-SL[a][j], z[j], s);
      z[a] = s * fast_rcp(SL[a][a]);
    }
  }

  // ================================================================== filter phase 1: chunk -> filtering element
  template <bool FIRST>
  static __device__ __forceinline__ void fold_step(Ctx& cx, const Lin& lin, long k, bool emit_pre,
                                                   real (&a)[R][D], real (&b)[R], real (&uf)[R][D],
                                                   real (&eta)[R], real (&z)[R][D], real (&gt)[R][D], int slot,
                                                   real* __restrict__ aggm, bool has_next) {
    stage_lin(cx, lin, k + 1, has_next);
    // ---- predict: A <- F A, b <- F b, T = tria([F Uf, QL])
    real t[R][D];
    {
      const real* M = publish<0>(cx, 0, a);
      mulF_rows<0>(cx, M, a);
      real bv[D];
      gather(cx, b, bv);
#pragma unroll
      for (int s = 0; s < R; ++s) {
        real acc = 0.0;
#pragma unroll
        for (int i = 0; i < Q1; ++i) acc = fma(cx.cf[s][i], pick(bv, cx.blk0[s] + i), acc);
        b[s] = acc;
      }
    }
    load_tq(cx, t);
    if constexpr (!FIRST) {
      real cc[R][D];
      const real* M2 = publish<d>(cx, 1, uf);
      mulF_rows<d>(cx, M2, cc);
      tpqrt<d, false>(cx, t, cc, nullptr, nullptr);
    }
    if (emit_pre) {
      constexpr int DD = D * D;
#pragma unroll
      for (int s = 0; s < R; ++s)
        if (cx.row[s] < D) {
          const int r = cx.row[s];
          aggm[DD + r] = b[s];
          aggm[2 * DD + D + r] = eta[s];
#pragma unroll
          for (int j = 0; j < D; ++j) {
            aggm[r * D + j] = a[s][j];
            aggm[DD + D + r * D + j] = (j <= r) ? t[s][j] : 0.0;
            aggm[2 * DD + 2 * D + r * D + j] = (j <= r) ? z[s][j] : 0.0;
          }
        }
    }
    // ---- update
    real SL[d][d];
    LinK lk;
    fetch_lin(cx, lin, k, lk);
    const real* MT = publish<0>(cx, 1, t);
    update(cx, lin, lk, MT, t, SL);
    // ---- G = SL^{-1} (H A) (columns over the owned rows' indices), zz = SL^{-1}(H b + c)
    real bv[D];
    gather(cx, b, bv);
    const real* MA = publish<0>(cx, 0, a);
    real g[d][R], rv[d], zz[d];
    H_times_vec(lin, lk, bv, rv);
    solveSL(SL, rv, zz);
#pragma unroll
    for (int s = 0; s < R; ++s) {
      real o[d], gg[d];
      H_times_col(lin, lk, MA, cx.rc[s], o);
      solveSL(SL, o, gg);
#pragma unroll
      for (int e = 0; e < d; ++e) g[e][s] = gg[e];
    }
    real Gf[d][D];
#pragma unroll
    for (int e = 0; e < d; ++e) gather(cx, g[e], Gf[e]);
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int e = 0; e < d; ++e) {
        const real kb = t[s][e];
        b[s] = fma(-kb, zz[e], b[s]);
        eta[s] = fma(-g[e][s], zz[e], eta[s]);
#pragma unroll
        for (int j = 0; j < D; ++j) a[s][j] = fma(-kb, Gf[e][j], a[s][j]);
      }
#pragma unroll
      for (int j = 0; j < D; ++j) uf[s][j] = (j < d) ? 0.0 : t[s][j];
    }
    // ---- pending columns of Z: G^T of this step goes to columns [D - (slot+1) d, D - slot d) of gt (flush_z absorbs)
#pragma unroll
    for (int sl = 0; sl < MZ; ++sl) {
#pragma unroll
      for (int s = 0; s < R; ++s) {
#pragma unroll
        for (int e = 0; e < d; ++e) gt[s][D - (sl + 1) * d + e] = (sl == slot) ? g[e][s] : gt[s][D - (sl + 1) * d + e];
      }
    }
  }
  // Z <- tria([Z | pending columns]); pending <- 0
  static __device__ __forceinline__ void flush_z(Ctx& cx, real (&z)[R][D], real (&gt)[R][D]) {
    tpqrt<D - MZ * d, false>(cx, z, gt, nullptr, nullptr);
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int j = 0; j < D; ++j) gt[s][j] = 0.0;
    }
  }

  static __device__ __forceinline__ void fold(Ctx& cx, long k0, long k1, const Lin& lin, real* __restrict__ agg,
                                              real* __restrict__ aggm) {
    real a[R][D], uf[R][D], z[R][D], b[R], eta[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      b[s] = 0.0;
      eta[s] = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        a[s][j] = (j == cx.row[s]) ? 1.0 : 0.0;
        uf[s][j] = 0.0;
        z[s][j] = 0.0;
      }
    }
    real gt[R][D];
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int j = 0; j < D; ++j) gt[s][j] = 0.0;
    }
    stage_lin(cx, lin, k0, true);
    fold_step<true>(cx, lin, k0, aggm && k0 == k1 - 1, a, b, uf, eta, z, gt, 0, aggm, k0 + 1 < k1);
    int slot = 1;
    for (long k = k0 + 1; k <= k1; ++k) {
      // one flush site: when MZ steps are pending, before the chunk's last step if it emits the pre-update element
      // (which must carry the complete Z), and after the last step
      if (slot == MZ || k == k1 || (k == k1 - 1 && aggm && slot != 0)) {
        flush_z(cx, z, gt);
        slot = 0;
      }
      if (k == k1) break;
      fold_step<false>(cx, lin, k, aggm && k == k1 - 1, a, b, uf, eta, z, gt, slot, aggm, k + 1 < k1);
      ++slot;
    }
    constexpr int DD = D * D;
#pragma unroll
    for (int s = 0; s < R; ++s)
      if (cx.row[s] < D) {
        const int r = cx.row[s];
        agg[DD + r] = b[s];
        agg[2 * DD + D + r] = eta[s];
#pragma unroll
        for (int j = 0; j < D; ++j) {
          agg[r * D + j] = a[s][j];
          agg[DD + D + r * D + j] = uf[s][j];
          agg[2 * DD + 2 * D + r * D + j] = (j <= r) ? z[s][j] : 0.0;
        }
      }
  }

  // ================================================================== filter phase 3: seeded square-root KF
  static __device__ __forceinline__ void store_row(real* __restrict__ p, const real (&x)[D]) {
    if constexpr (D % 2 == 0) {
      real2* p2 = reinterpret_cast<real2*>(p);
#pragma unroll
      for (int j = 0; j < D / 2; ++j) p2[j] = make_real2(x[2 * j], x[2 * j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < D; ++j) p[j] = x[j];
    }
  }
  struct Stats {
    real nll, s1, s2;
  };
  // J0 = first non-zero column of the incoming factor (0 for a chunk's first step, d afterwards)
  template <int J0>
  static __device__ __forceinline__ void scan_step(Ctx& cx, const Lin& lin, long k, real (&m)[R], real (&uf)[R][D],
                                                   real* __restrict__ kern, Stats& st, real* __restrict__ fmeans,
                                                   real* __restrict__ fchols, bool has_next) {
    stage_lin(cx, lin, k + 1, has_next);
    // ---- predict + backward kernel: [[F Uf, QL],[Uf, 0]] -> [[T, 0],[Phi21, Phi22~]]
    real t[R][D], cc[R][D], e[R][D];
    {
      const real* M = publish<J0>(cx, 0, uf);
      mulF_rows<J0>(cx, M, cc);
    }
    load_tq(cx, t);
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int j = 0; j < D; ++j) e[s][j] = 0.0;
    }
    tpqrt<J0, true>(cx, t, cc, e, uf);
    // ---- E rows: e <- e T^{-1}
    const real* MT = publish<0>(cx, 1, t);
    {
      real dinv[R], inv[D];
#pragma unroll
      for (int s = 0; s < R; ++s) dinv[s] = fast_rcp(pick(t[s], cx.rc[s]));
      gather(cx, dinv, inv);
#pragma unroll
      for (int j = D - 1; j >= 0; --j) {
        real acc[R];
#pragma unroll
        for (int s = 0; s < R; ++s) acc[s] = e[s][j];
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
          const real tv = MT[i * LDM + j];
#pragma unroll
          for (int s = 0; s < R; ++s) acc[s] = fma(-e[s][i], tv, acc[s]);
        }
#pragma unroll
        for (int s = 0; s < R; ++s) e[s][j] = acc[s] * inv[j];
      }
    }
    // ---- means: mp = F m ; g = m - E mp
    real mv[D], g[R];
    gather(cx, m, mv);
    mulF_vec(mv);
#pragma unroll
    for (int s = 0; s < R; ++s) {
      real acc = m[s];
#pragma unroll
      for (int i = 0; i < D; ++i) acc = fma(-e[s][i], mv[i], acc);
      g[s] = acc;
    }
    // ---- store the step's backward kernel (g | E | Phi22~): the noise factor stays UNtriangularised -- the smoother
    // only needs Phi22~ Phi22~^T and triangularises [E L | Phi22~] in one go (D - J0 - 1 pivots less per step here)
    {
      real* kp = kern + k * NE;
#pragma unroll
      for (int s = 0; s < R; ++s)
        if (cx.row[s] < D) kp[cx.row[s]] = g[s];
      store_rows_pm<0>(cx, kp + D, e);
      store_rows_pm<(J0 / PW) * PW>(cx, kp + D + D * D, uf);
    }
    // ---- measurement update
    real SL[d][d], y[d], zz[d];
    LinK lk;
    fetch_lin(cx, lin, k, lk);
    update(cx, lin, lk, MT, t, SL);
    H_times_vec(lin, lk, mv, y);
    solveSL(SL, y, zz);
#pragma unroll
    for (int s = 0; s < R; ++s) {
      real acc = pick(mv, cx.rc[s]);
#pragma unroll
      for (int a = 0; a < d; ++a) acc = fma(-t[s][a], zz[a], acc);
      m[s] = acc;
#pragma unroll
      for (int j = 0; j < D; ++j) uf[s][j] = (j < d) ? 0.0 : t[s][j];
    }
    // ---- innovation statistics (replicated)
    real q2 = 0.0, lg = 0.0;
#pragma unroll
    for (int a = 0; a < d; ++a) {
      q2 = fma(zz[a], zz[a], q2);
      lg += log(fabs(SL[a][a]));
    }
    st.nll += 0.5 * q2 + lg + 0.5 * d * LOG_2PI;
    st.s2 += q2;
    real wv[d], ww = 0.0;
#pragma unroll
    for (int a = d - 1; a >= 0; --a) {
      real acc = y[a];
#pragma unroll
      for (int e2 = a + 1; e2 < d; ++e2) acc = fma(-SL[e2][a], wv[e2], acc);
      wv[a] = acc * fast_rcp(SL[a][a]);
      ww = fma(wv[a], wv[a], ww);
    }
    st.s1 += ww;
    if (fmeans) {
#pragma unroll
      for (int s = 0; s < R; ++s)
        if (cx.row[s] < D) {
          fmeans[(k + 1) * D + cx.row[s]] = m[s];
#pragma unroll
          for (int j = 0; j < D; ++j) fchols[((k + 1) * D + cx.row[s]) * D + j] = uf[s][j];
        }
    }
  }

  static __device__ __forceinline__ void scan(Ctx& cx, long k0, long k1, const Lin& lin,
                                              const real* __restrict__ state_in, real* __restrict__ kern,
                                              real* __restrict__ state_end, real* __restrict__ part,
                                              real* __restrict__ fmeans, real* __restrict__ fchols) {
    real m[R], uf[R][D];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const bool ok = cx.row[s] < D;
      m[s] = ok ? state_in[cx.rc[s]] : 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) uf[s][j] = ok ? state_in[D + cx.rc[s] * D + j] : 0.0;
    }
    Stats st = {0.0, 0.0, 0.0};
    stage_lin(cx, lin, k0, true);
    scan_step<0>(cx, lin, k0, m, uf, kern, st, fmeans, fchols, k0 + 1 < k1);
    for (long k = k0 + 1; k < k1; ++k) scan_step<d>(cx, lin, k, m, uf, kern, st, fmeans, fchols, k + 1 < k1);
    // filtered end state with a lower-triangular factor
    real le[R][D];
    tria_rows<d>(cx, uf, le);
#pragma unroll
    for (int s = 0; s < R; ++s)
      if (cx.row[s] < D) {
        state_end[cx.row[s]] = m[s];
#pragma unroll
        for (int j = 0; j < D; ++j) state_end[D + cx.row[s] * D + j] = le[s][j];
      }
    if (cx.l == 0) {
      part[0] = st.nll;
      part[1] = st.s1;
      part[2] = st.s2;
    }
  }

  // ================================================================== smoother phase 3: seeded square-root RTS
  // a step's backward kernel (g | E | Dk), row-contiguous per lane: (g, E rows) and (Dk rows) are loaded separately so
  // that only (g, E) -- needed first -- is held one step ahead in registers; Dk is loaded at the top of its own step
  // from L2, where a prefetch issued one step earlier has put it (holding all 2D^2+D doubles of the next step in
  // registers spilled, and the spill stores then waited for the loads: ncu STL on long_sb, 19 % of the samples)
  // Matrices of the backward kernels (private to scan -> smooth) are stored PIECE-major: element (r, c) of a D x D
  // block lives at ((c / PW) * D + r) * PW + c % PW with PW = 2 doubles (one 128-bit access) for even D.  The G lanes
  // of a group then touch G consecutive 16-byte pieces per instruction (one 64-byte segment per chunk) instead of G
  // pieces 8 D bytes apart: 8 instead of ~20 L1 tag look-ups and half the L2 sectors per warp access (ncu, r01).
  static constexpr int PW = (D % 2 == 0) ? 2 : 1;
  template <int JS = 0>  // columns [JS, D) only (JS a multiple of PW)
  static __device__ __forceinline__ void load_rows(const Ctx& cx, const real* __restrict__ base, real (&x)[R][D]) {
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int rc = cx.rc[s];
#pragma unroll
      for (int j = 0; j < JS; ++j) x[s][j] = 0.0;
      if constexpr (PW == 2) {
        const real2* pe = reinterpret_cast<const real2*>(base) + rc;
#pragma unroll
        for (int j = JS / 2; j < D / 2; ++j) {
          const real2 a = pe[j * D];
          x[s][2 * j] = a.x;
          x[s][2 * j + 1] = a.y;
        }
      } else {
#pragma unroll
        for (int j = JS; j < D; ++j) x[s][j] = base[j * D + rc];
      }
    }
  }
  template <int JS>
  static __device__ __forceinline__ void store_rows_pm(const Ctx& cx, real* __restrict__ base,
                                                       const real (&x)[R][D]) {
#pragma unroll
    for (int s = 0; s < R; ++s)
      if (cx.row[s] < D) {
        const int r = cx.row[s];
        if constexpr (PW == 2) {
          real2* pe = reinterpret_cast<real2*>(base) + r;
#pragma unroll
          for (int j = JS / 2; j < D / 2; ++j) pe[j * D] = make_real2(x[s][2 * j], x[s][2 * j + 1]);
        } else {
#pragma unroll
          for (int j = JS; j < D; ++j) base[j * D + r] = x[s][j];
        }
      }
  }
  // L2 prefetch of one step's backward kernel (NE contiguous doubles) and the previous mean of that row, spread over
  // the group's lanes in 128-byte strides
  static __device__ __forceinline__ void prefetch_step(const Ctx& cx, const real* __restrict__ kp,
                                                       const real* __restrict__ mrow) {
#pragma unroll
    for (int o = 0; o < NE; o += 16 * G) {
      const int off = o + 16 * cx.l;
      if (off < NE) prefetch_l2(kp + off);
    }
    if (cx.l == 0) {
      prefetch_l2(kp + NE - 1);
      prefetch_l2(mrow);
      prefetch_l2(mrow + D - 1);
    }
  }
  static __device__ __forceinline__ real emit(const Ctx& cx, long t, const real (&m)[R], const real (&l)[R][D],
                                                real cscale, const real (&old)[R], real* __restrict__ means,
                                                real* __restrict__ chols) {
    real bad = 0.0;
#pragma unroll
    for (int s = 0; s < R; ++s)
      if (cx.row[s] < D) {
        const int r = cx.row[s];
        const bool close = fabs(old[s] - m[s]) <= (1e-8 + 1e-13 * fabs(m[s]));
        bad += close ? 0.0 : 1.0;
        means[t * D + r] = m[s];
        if (chols) {
          real rowv[D];
#pragma unroll
          for (int j = 0; j < D; ++j) rowv[j] = (j <= r) ? cscale * l[s][j] : 0.0;
          store_row(chols + (t * D + r) * D, rowv);
        }
      }
    return bad;
  }
  static __device__ __forceinline__ unsigned stg_bar(const Ctx& cx) {
    return (unsigned)__cvta_generic_to_shared(cx.stg + NE);
  }
  // leader lane: arm the group's mbarrier with the byte count and start the bulk copy of one step's kernel
  static __device__ __forceinline__ void stage_issue(const Ctx& cx, const real* __restrict__ src) {
    if (cx.l == 0) {
      const unsigned bar = stg_bar(cx), dst = (unsigned)__cvta_generic_to_shared(cx.stg);
      constexpr unsigned BYTES = NE * sizeof(real);
      // (write-after-read on the block: the group's reads completed before the __syncwarp that precedes this call --
      // the same consumer-release ordering TMA pipelines rely on; a fence.proxy.async here cost ~2 us per call)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BYTES) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(src), "r"(BYTES), "r"(bar)
                   : "memory");
    }
  }
  // NON-blocking poll (mbarrier.test_wait): the groups of a warp wait on different barriers, and the potentially
  // blocking try_wait form suspends each diverged group in turn (measured: 125 us per step instead of 4).  The copy was
  // issued a whole step earlier, so the first poll normally succeeds and the warp does not diverge at all.
  static __device__ __forceinline__ void stage_wait(Ctx& cx) {
    const unsigned bar = stg_bar(cx);
    unsigned done = 0;
    do {
      asm volatile(
          "{\n"
          ".reg .pred P1;\n"
          "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
          "selp.u32 %0, 1, 0, P1;\n"
          "}"
          : "=r"(done)
          : "r"(bar), "r"(cx.stg_phase)
          : "memory");
    } while (!done);
    cx.stg_phase ^= 1u;
  }
  static __device__ __forceinline__ void stage_init(Ctx& cx, real* block) {
    cx.stg = block;
    cx.stg_phase = 0;
    if (cx.l == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(stg_bar(cx)), "r"(1) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cx.sync();
  }

  template <int JS>
  static __device__ __forceinline__ void smooth_step(Ctx& cx, long k, bool has_prev, bool emit_t0,
                                                     const real* qLinvdiag, const real* qL,
                                                     const real* __restrict__ kern, real cscale,
                                                     real* __restrict__ means, real* __restrict__ chols,
                                                     real (&m)[R], real (&l)[R][D], real& obj, real& bad) {
    real g[R], e[R][D], ph[R][D], old[R];
    if (cx.stg) {
      stage_wait(cx);  // this step's kernel has landed in the group's staging block
      const real* kp = cx.stg;
#pragma unroll
      for (int s = 0; s < R; ++s) {
        g[s] = kp[cx.rc[s]];
        old[s] = (k > 0 || emit_t0) ? means[k * D + cx.rc[s]] : 0.0;
      }
      load_rows<0>(cx, kp + D, e);
      load_rows<JS>(cx, kp + D + D * D, ph);
      cx.sync();  // every lane of the group has its rows: the block may be overwritten
      if (has_prev) {
        stage_issue(cx, kern + (k - 1) * NE);
        if (cx.l == 0) {
          prefetch_l2(means + (k - 1) * D);
          prefetch_l2(means + (k - 1) * D + D - 1);
        }
      }
    } else {
      const real* kp = kern + k * NE;
#pragma unroll
      for (int s = 0; s < R; ++s) {
        g[s] = kp[cx.rc[s]];
        old[s] = (k > 0 || emit_t0) ? means[k * D + cx.rc[s]] : 0.0;
      }
      load_rows<0>(cx, kp + D, e);
      load_rows<JS>(cx, kp + D + D * D, ph);
      if (has_prev) prefetch_step(cx, kern + (k - 1) * NE, means + (k - 1) * D);
    }
    const real* ML = publish<0>(cx, 0, l);
    real mv[D];
    gather(cx, m, mv);
    real mn[R], cd[R][D];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      mn[s] = g[s];
#pragma unroll
      for (int j = 0; j < D; ++j) cd[s][j] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < D; ++i) {
#pragma unroll
      for (int s = 0; s < R; ++s) mn[s] = fma(e[s][i], mv[i], mn[s]);
#pragma unroll
      for (int j = 0; j <= i; ++j) {
        const real lv = ML[i * LDM + j];
#pragma unroll
        for (int s = 0; s < R; ++s) cd[s][j] = fma(e[s][i], lv, cd[s][j]);
      }
    }
    tria2_step<0, JS>(cx, cd, ph);
    // objective increment |QL^{-1}(m_k - F m_{k+1})|^2 (replicated)
    real fm[D], rr[D], dr[R];
#pragma unroll
    for (int i = 0; i < D; ++i) fm[i] = mv[i];
    mulF_vec(fm);
#pragma unroll
    for (int s = 0; s < R; ++s) dr[s] = mn[s] - pick(fm, cx.rc[s]);
    gather(cx, dr, rr);
#pragma unroll
    for (int b = 0; b < d; ++b) {
#pragma unroll
      for (int i = 0; i < Q1; ++i) {
        real acc = rr[b * Q1 + i];
#pragma unroll
        for (int j = 0; j < i; ++j) acc = fma(-qL[i * Q1 + j], rr[b * Q1 + j], acc);
        acc *= qLinvdiag[i];
        rr[b * Q1 + i] = acc;
        obj = fma(acc, acc, obj);
      }
    }
#pragma unroll
    for (int s = 0; s < R; ++s) {
      m[s] = mn[s];
#pragma unroll
      for (int j = 0; j < D; ++j) l[s][j] = (j <= cx.rc[s]) ? cd[s][j] : 0.0;
    }
    if (k > 0 || emit_t0) bad += emit(cx, k, m, l, cscale, old, means, chols);
  }
  static __device__ __forceinline__ void smooth(Ctx& cx, long k0, long k1, bool last, bool emit_t0,
                                                const real* qLinvdiag, const real* qL,
                                                const real* __restrict__ seed, const real* __restrict__ kern,
                                                real cscale, real* __restrict__ means,
                                                real* __restrict__ chols, real* __restrict__ part) {
    real m[R], l[R][D], old[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const bool ok = cx.row[s] < D;
      m[s] = ok ? seed[cx.rc[s]] : 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) l[s][j] = (ok && j <= cx.rc[s]) ? seed[D + cx.rc[s] * D + j] : 0.0;
    }
    real obj = 0.0, bad = 0.0;
    if (last) {
#pragma unroll
      for (int s = 0; s < R; ++s) old[s] = means[k1 * D + cx.rc[s]];
      bad += emit(cx, k1, m, l, cscale, old, means, chols);
    }
    // Global loads of a step are issued at its top and hit L2: the whole kernel of step k-1 (and the previous mean of
    // row k-1) is prefetched into L2 while step k computes.  (Holding the next kernel in registers spilled; mixing
    // DRAM-latency loads for step k-1 with L2 hits for step k made every consumer wait for the slowest load, because
    // the few load scoreboards are shared -- ncu: long_scoreboard on the first shuffle of Dk, 33 % of the samples.)
    if (cx.stg)
      stage_issue(cx, kern + (k1 - 1) * NE);
    else
      prefetch_step(cx, kern + (k1 - 1) * NE, means + (k1 - 1) * D);
    // the noise factor of a step has D - d columns, except for the chunk's first step (full incoming factor)
    for (long k = k1 - 1; k > k0; --k)
      smooth_step<(d / PW) * PW>(cx, k, true, emit_t0, qLinvdiag, qL, kern, cscale, means, chols, m, l, obj, bad);
    smooth_step<0>(cx, k0, false, emit_t0, qLinvdiag, qL, kern, cscale, means, chols, m, l, obj, bad);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) bad += __shfl_xor_sync(cx.mask, bad, o, G);
    if (cx.l == 0) {
      part[0] = obj;
      part[1] = bad;
    }
  }
};

}  // namespace POF_NS
