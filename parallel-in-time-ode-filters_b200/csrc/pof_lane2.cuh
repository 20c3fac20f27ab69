// Leaf level, performance path v2: G lanes of a warp cooperate on ONE time-chunk and every lane owns R = 2 ROWS of each
// D-row matrix (row = lane + s*G, s = 0,1), G = pow2ceil(ceil(D/2)): D = 8 -> 4 lanes per chunk, 8 chunks per warp.
//
// Why two rows per lane (measured on B200 with the one-row kernels of pof_lane.cuh, profiles/r01_*): every
// Householder pivot has to deliver its K+1 pivot-row entries to every lane, 2 crossbar wavefronts per double per warp
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

#include "pof_small.cuh"

namespace pof {

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
  static constexpr double LOG_2PI = 1.8378770664093454835606594728112;
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

  struct Lin {
    const double* __restrict__ H;
    const double* __restrict__ c;
    const double* __restrict__ Jc;
    double s0, s1;
  };

  struct Ctx {
    int l;             // lane within the group
    int row[R];        // rows owned (may be >= D: idle slot)
    int rc[R];         // clamped row index for addressing
    int rb[R], blk0[R];
    unsigned mask;
    double* mat;       // 2 exchange matrices
    double* vec;       // 2 gather vectors
    double* tq;        // this lane's rows of QL, lane-minor: tq[(s*D + j)*G] (conflict-free across the group)
    double* lbuf;      // 2 x NJP doubles: compact linearisations staged by cp.async (global -> shared, no registers)
    int vflip;
    double cf[R][Q1];  // Pascal coefficients of the owned rows of F
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
  };

  static __device__ __forceinline__ void init_ctx(Ctx& c, double* sm_group, const double* qL) {
    const int lane = threadIdx.x & 31;
    c.l = lane % G;
    c.mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
    c.mat = sm_group;
    c.vec = sm_group + 2 * D * LDM;
    c.tq = c.vec + 2 * VEC + c.l;
    c.lbuf = c.vec + 2 * VEC + R * G * D;
    c.vflip = 0;
#pragma unroll
    for (int s = 0; s < R; ++s) {
      c.row[s] = c.l + s * G;
      c.rc[s] = (c.row[s] < D) ? c.row[s] : 0;
      c.rb[s] = c.rc[s] % Q1;
      c.blk0[s] = c.rc[s] - c.rb[s];
#pragma unroll
      for (int i = 0; i < Q1; ++i) {
        double v = 0.0;
#pragma unroll
        for (int b = 0; b < Q1; ++b)
          if (c.rb[s] == b && i >= b) v = binom(q - b, i - b);
        c.cf[s][i] = v;
      }
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double v = 0.0;
#pragma unroll
        for (int b = 0; b < Q1; ++b)
          if (c.rb[s] == b && (j / Q1) * Q1 == c.blk0[s] && (j % Q1) <= b) v = qL[b * Q1 + (j % Q1)];
        c.tq[(s * D + j) * G] = (c.row[s] < D) ? v : 0.0;
      }
    }
    c.sync();
  }

  // ---------------------------------------------------------------------------------------------- small helpers
  static __device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
  }
  static __device__ __forceinline__ double fast_rsqrt(double x) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double hx = 0.5 * x;
    double e = fma(-hx * r, r, 0.5);
    r = fma(r, e, r);
    e = fma(-hx * r, r, 0.5);
    return fma(r, e, r);
  }
  // sum_j a[j]*b[j] with several independent accumulators: the leaf recursions are bound by dependent-issue latency
  // ("wait" stalls with 2 warps per scheduler), so every long FMA chain is split and tree-added
  template <int n>
  static __device__ __forceinline__ double dotn(const double* a, const double* b) {
    if constexpr (n <= 0) {
      return 0.0;
    } else if constexpr (n < 4) {
      double s = a[0] * b[0];
#pragma unroll
      for (int j = 1; j < n; ++j) s = fma(a[j], b[j], s);
      return s;
    } else if constexpr (n < 10) {
      double s0 = a[0] * b[0], s1 = a[1] * b[1];
#pragma unroll
      for (int j = 2; j + 1 < n; j += 2) {
        s0 = fma(a[j], b[j], s0);
        s1 = fma(a[j + 1], b[j + 1], s1);
      }
      if constexpr (n % 2) s0 = fma(a[n - 1], b[n - 1], s0);
      return s0 + s1;
    } else {
      double s0 = a[0] * b[0], s1 = a[1] * b[1], s2 = a[2] * b[2], s3 = a[3] * b[3];
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
    double s, tp, beta;
  };
  // Householder for the row (alpha, x[0..n)):  H = I - tp v v^T, v = (s, x), H (alpha, x)^T = (beta, 0)
  template <int n>
  static __device__ __forceinline__ HH house(double alpha, const double* x) {
    const double sigma = dotn<n>(x, x);
    const double nrm2 = fma(alpha, alpha, sigma);
    // rsqrt.approx.ftz flushes subnormal inputs to zero (-> inf -> NaN in the Newton step): a row whose squared norm
    // is below 2^-1000 is treated as already reduced (identity reflector), as a zero tail is
    const bool nz = sigma > 0.0 && nrm2 > 0x1p-1000;
    const double rn = fast_rsqrt(nrm2);
    const double nrm = nrm2 * rn;
    const double beta = (alpha >= 0.0) ? -nrm : nrm;
    const double s = alpha - beta;
    HH h;
    h.beta = nz ? beta : alpha;
    h.s = nz ? s : 0.0;
    h.tp = nz ? rn * fast_rcp(fabs(s)) : 0.0;
    return h;
  }
  static __device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
  }
  static __device__ __forceinline__ double bshfl(const Ctx& c, double x, int src) {
    return __shfl_sync(c.mask, x, ((threadIdx.x & 31) / G) * G + src);
  }
  // out[row] = x[s] of the lane owning `row`
  static __device__ __forceinline__ void gather(Ctx& c, const double (&x)[R], double (&out)[D]) {
    double* v = c.vec + c.vflip * VEC;
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
  static __device__ __forceinline__ const double* publish(Ctx& c, int which, const double (&x)[R][D]) {
    double* M = c.mat + which * D * LDM;
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
  static __device__ __forceinline__ void mulF_rows(const Ctx& c, const double* M, double (&y)[R][D]) {
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int j = J0; j < D; ++j) y[s][j] = 0.0;
#pragma unroll
      for (int i = 0; i < Q1; ++i) {
        const double* rowp = M + (c.blk0[s] + i) * LDM;
#pragma unroll
        for (int j = J0; j < D; ++j) y[s][j] = fma(c.cf[s][i], rowp[j], y[s][j]);
      }
    }
  }
  static __device__ __forceinline__ void mulF_vec(double (&m)[D]) {
#pragma unroll
    for (int b = 0; b < d; ++b) {
#pragma unroll
      for (int i = 0; i < Q1; ++i) {
#pragma unroll
        for (int j = i + 1; j < Q1; ++j) m[b * Q1 + i] = fma(binom(q - i, j - i), m[b * Q1 + j], m[b * Q1 + i]);
      }
    }
  }
  static __device__ __forceinline__ double pick(const double (&x)[D], int idx) {
    double v = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i)
      if (i == idx) v = x[i];
    return v;
  }
  static __device__ __forceinline__ void load_tq(const Ctx& c, double (&t)[R][D]) {
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
  static __device__ __forceinline__ void tp_step2(Ctx& cx, double (&t)[R][D], double (&c)[R][D], double (*pt)[D],
                                                  double (*pc)[D]) {
    if constexpr (I < D) {
      constexpr int K = D - J0;
      constexpr int so = I / G;
      const int lo = I % G;
      double piv[K + 1];
      piv[0] = bshfl(cx, t[so][I], lo);
#pragma unroll
      for (int j = 0; j < K; ++j) piv[1 + j] = bshfl(cx, c[so][J0 + j], lo);
      const HH h = house<K>(piv[0], piv + 1);
      row_update<0, I, J0, K>(cx, h, piv, t, c);
      row_update<1, I, J0, K>(cx, h, piv, t, c);
      if constexpr (PASS) {
#pragma unroll
        for (int s = 0; s < R; ++s) {
          double u = fma(h.s, pt[s][I], dotn<K>(&pc[s][J0], piv + 1));
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
  static __device__ __forceinline__ void row_update(const Ctx& cx, const HH& h, const double* piv, double (&t)[R][D],
                                                    double (&c)[R][D]) {
    if constexpr (S * G + G - 1 >= I) {  // some row of this slot is at or below the pivot
      double w = fma(h.s, t[S][I], dotn<K>(&c[S][J0], piv + 1));  // the dot does not wait for the reflector
      w = (cx.row[S] >= I) ? w * h.tp : 0.0;
      t[S][I] = (cx.row[S] == I) ? h.beta : fma(-w, h.s, t[S][I]);
#pragma unroll
      for (int j = 0; j < K; ++j) c[S][J0 + j] = fma(-w, piv[1 + j], c[S][J0 + j]);
    }
  }
  template <int J0, bool PASS>
  static __device__ __forceinline__ void tpqrt(Ctx& cx, double (&t)[R][D], double (&c)[R][D], double (*pt)[D],
                                               double (*pc)[D]) {
    tp_step2<0, J0, PASS>(cx, t, c, pt, pc);
  }

  // plain right-Householder lower-triangularisation of the D x (D - J0) matrix held in columns [J0, D) of x; the
  // result is written as a lower-triangular D x D factor into columns [0, D - J0) (pivot I uses column J0 + I)
  template <int S, int I, int J0>
  static __device__ __forceinline__ void tria_update(const Ctx& cx, const HH& h, const double* piv,
                                                     double (&x)[R][D]) {
    if constexpr (S * G + G - 1 >= I) {
      constexpr int n = D - J0 - I;
      double w = fma(h.s, x[S][J0 + I], dotn<n - 1>(&x[S][J0 + I + 1], piv + 1));
      w = (cx.row[S] >= I) ? w * h.tp : 0.0;
      x[S][J0 + I] = (cx.row[S] == I) ? h.beta : fma(-w, h.s, x[S][J0 + I]);
#pragma unroll
      for (int j = 1; j < n; ++j) x[S][J0 + I + j] = fma(-w, piv[j], x[S][J0 + I + j]);
    }
  }
  template <int I, int J0>
  static __device__ __forceinline__ void tria_step(Ctx& cx, double (&x)[R][D]) {
    if constexpr (I + 1 < D - J0) {
      constexpr int n = D - J0 - I;
      constexpr int so = I / G;
      const int lo = I % G;
      double piv[n];
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
  static __device__ __forceinline__ void tria_rows(Ctx& cx, double (&x)[R][D], double (&out)[R][D]) {
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
  static __device__ __forceinline__ void tria2_update(const Ctx& cx, const HH& h, const double* piv, double (&x)[R][D],
                                                      double (&y)[R][D]) {
    if constexpr (S * G + G - 1 >= I) {
      constexpr int n1 = D - I, n2 = D - JS;
      double w = fma(h.s, x[S][I], dotn<n1 - 1>(&x[S][I + 1], piv + 1) + dotn<n2>(&y[S][JS], piv + n1));
      w = (cx.row[S] >= I) ? w * h.tp : 0.0;
      x[S][I] = (cx.row[S] == I) ? h.beta : fma(-w, h.s, x[S][I]);
#pragma unroll
      for (int j = 1; j < n1; ++j) x[S][I + j] = fma(-w, piv[j], x[S][I + j]);
#pragma unroll
      for (int j = 0; j < n2; ++j) y[S][JS + j] = fma(-w, piv[n1 + j], y[S][JS + j]);
    }
  }
  template <int I, int JS>
  static __device__ __forceinline__ void tria2_step(Ctx& cx, double (&x)[R][D], double (&y)[R][D]) {
    if constexpr (I < D) {
      constexpr int n1 = D - I, n2 = D - JS;
      constexpr int so = I / G;
      const int lo = I % G;
      double piv[n1 + n2];
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
    double J[d][d], c[d];
    const double* Hd;  // dense H of this step, or null
  };
  static __device__ __forceinline__ void load_lin(const Lin& L, long k, LinK& o) {
    if (L.Jc) {
      const double* p = L.Jc + k * (d * d + d);
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
      const double* p = L.Jc + k * (d * d + d);
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
      const double* src = L.Jc + k * NJ;
      double* dst = cx.lbuf + (k & 1) * NJP;
#pragma unroll
      for (int i0 = 0; i0 < NJ / CPD; i0 += G) {
        const int i = i0 + cx.l;
        if (i < NJ / CPD) {
          const unsigned sa = (unsigned)__cvta_generic_to_shared(dst + i * CPD);
          if constexpr (CPD == 2)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(src + i * CPD) : "memory");
          else
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(src + i * CPD) : "memory");
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
      const double* p = cx.lbuf + (k & 1) * NJP;
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
  static __device__ __forceinline__ void H_times_col(const Lin& L, const LinK& lk, const double* M, int col,
                                                     double (&out)[d]) {
    if (lk.Hd == nullptr) {
#pragma unroll
      for (int a = 0; a < d; ++a) {
        double s = L.s1 * M[(a * Q1 + 1) * LDM + col];
#pragma unroll
        for (int b = 0; b < d; ++b) s = fma(-L.s0 * lk.J[a][b], M[(b * Q1) * LDM + col], s);
        out[a] = s;
      }
    } else {
#pragma unroll
      for (int a = 0; a < d; ++a) out[a] = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const double v = M[i * LDM + col];
#pragma unroll
        for (int a = 0; a < d; ++a) out[a] = fma(__ldg(lk.Hd + a * D + i), v, out[a]);
      }
    }
  }
  // H v + c for a replicated vector v
  static __device__ __forceinline__ void H_times_vec(const Lin& L, const LinK& lk, const double (&v)[D],
                                                     double (&out)[d]) {
    if (lk.Hd == nullptr) {
#pragma unroll
      for (int a = 0; a < d; ++a) {
        double s = fma(L.s1, v[a * Q1 + 1], lk.c[a]);
#pragma unroll
        for (int b = 0; b < d; ++b) s = fma(-L.s0 * lk.J[a][b], v[b * Q1], s);
        out[a] = s;
      }
    } else {
#pragma unroll
      for (int a = 0; a < d; ++a) {
        double s = lk.c[a];
#pragma unroll
        for (int i = 0; i < D; ++i) s = fma(__ldg(lk.Hd + a * D + i), v[i], s);
        out[a] = s;
      }
    }
  }

  // ---------------------------------------------------------------------------------------------- measurement update
  template <int A>
  static __device__ __forceinline__ void update_pivot(double (&t)[R][D], double (&W)[d][D]) {
    if constexpr (A < d) {
      const HH h = house<D - A - 1>(W[A][A], &W[A][A + 1]);
#pragma unroll
      for (int s = 0; s < R; ++s) {
        double w = fma(h.s, t[s][A], dotn<D - A - 1>(&t[s][A + 1], &W[A][A + 1]));
        w *= h.tp;
        t[s][A] = fma(-w, h.s, t[s][A]);
#pragma unroll
        for (int j = A + 1; j < D; ++j) t[s][j] = fma(-w, W[A][j], t[s][j]);
      }
#pragma unroll
      for (int a2 = A + 1; a2 < d; ++a2) {
        double u = fma(h.s, W[a2][A], dotn<D - A - 1>(&W[a2][A + 1], &W[A][A + 1]));
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
  static __device__ __forceinline__ void update(Ctx& cx, const Lin& L, const LinK& lk, const double* MT,
                                                double (&t)[R][D], double (&SL)[d][d]) {
    double wc[d][R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      double o[d];
      H_times_col(L, lk, MT, cx.rc[s], o);
#pragma unroll
      for (int a = 0; a < d; ++a) wc[a][s] = o[a];
    }
    double W[d][D];
#pragma unroll
    for (int a = 0; a < d; ++a) gather(cx, wc[a], W[a]);
    update_pivot<0>(t, W);
#pragma unroll
    for (int a = 0; a < d; ++a) {
#pragma unroll
      for (int e = 0; e < d; ++e) SL[a][e] = (e <= a) ? W[a][e] : 0.0;
    }
  }
  static __device__ __forceinline__ void solveSL(const double (&SL)[d][d], const double (&y)[d], double (&z)[d]) {
#pragma unroll
    for (int a = 0; a < d; ++a) {
      double s = y[a];
#pragma unroll
      for (int j = 0; j < a; ++j) s = fma(-SL[a][j], z[j], s);
      z[a] = s * fast_rcp(SL[a][a]);
    }
  }

  // ================================================================== filter phase 1: chunk -> filtering element
  template <bool FIRST>
  static __device__ __forceinline__ void fold_step(Ctx& cx, const Lin& lin, long k, bool emit_pre,
                                                   double (&a)[R][D], double (&b)[R], double (&uf)[R][D],
                                                   double (&eta)[R], double (&z)[R][D], double (&gt)[R][D], int slot,
                                                   double* __restrict__ aggm, bool has_next) {
    stage_lin(cx, lin, k + 1, has_next);
    // ---- predict: A <- F A, b <- F b, T = tria([F Uf, QL])
    double t[R][D];
    {
      const double* M = publish<0>(cx, 0, a);
      mulF_rows<0>(cx, M, a);
      double bv[D];
      gather(cx, b, bv);
#pragma unroll
      for (int s = 0; s < R; ++s) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < Q1; ++i) acc = fma(cx.cf[s][i], pick(bv, cx.blk0[s] + i), acc);
        b[s] = acc;
      }
    }
    load_tq(cx, t);
    if constexpr (!FIRST) {
      double cc[R][D];
      const double* M2 = publish<d>(cx, 1, uf);
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
    double SL[d][d];
    LinK lk;
    fetch_lin(cx, lin, k, lk);
    const double* MT = publish<0>(cx, 1, t);
    update(cx, lin, lk, MT, t, SL);
    // ---- G = SL^{-1} (H A) (columns over the owned rows' indices), zz = SL^{-1}(H b + c)
    double bv[D];
    gather(cx, b, bv);
    const double* MA = publish<0>(cx, 0, a);
    double g[d][R], rv[d], zz[d];
    H_times_vec(lin, lk, bv, rv);
    solveSL(SL, rv, zz);
#pragma unroll
    for (int s = 0; s < R; ++s) {
      double o[d], gg[d];
      H_times_col(lin, lk, MA, cx.rc[s], o);
      solveSL(SL, o, gg);
#pragma unroll
      for (int e = 0; e < d; ++e) g[e][s] = gg[e];
    }
    double Gf[d][D];
#pragma unroll
    for (int e = 0; e < d; ++e) gather(cx, g[e], Gf[e]);
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int e = 0; e < d; ++e) {
        const double kb = t[s][e];
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
  static __device__ __forceinline__ void flush_z(Ctx& cx, double (&z)[R][D], double (&gt)[R][D]) {
    tpqrt<D - MZ * d, false>(cx, z, gt, nullptr, nullptr);
#pragma unroll
    for (int s = 0; s < R; ++s) {
#pragma unroll
      for (int j = 0; j < D; ++j) gt[s][j] = 0.0;
    }
  }

  static __device__ __forceinline__ void fold(Ctx& cx, long k0, long k1, const Lin& lin, double* __restrict__ agg,
                                              double* __restrict__ aggm) {
    double a[R][D], uf[R][D], z[R][D], b[R], eta[R];
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
    double gt[R][D];
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
  static __device__ __forceinline__ void store_row(double* __restrict__ p, const double (&x)[D]) {
    if constexpr (D % 2 == 0) {
      double2* p2 = reinterpret_cast<double2*>(p);
#pragma unroll
      for (int j = 0; j < D / 2; ++j) p2[j] = make_double2(x[2 * j], x[2 * j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < D; ++j) p[j] = x[j];
    }
  }
  struct Stats {
    double nll, s1, s2;
  };
  // J0 = first non-zero column of the incoming factor (0 for a chunk's first step, d afterwards)
  template <int J0>
  static __device__ __forceinline__ void scan_step(Ctx& cx, const Lin& lin, long k, double (&m)[R], double (&uf)[R][D],
                                                   double* __restrict__ kern, Stats& st, double* __restrict__ fmeans,
                                                   double* __restrict__ fchols, bool has_next) {
    stage_lin(cx, lin, k + 1, has_next);
    // ---- predict + backward kernel: [[F Uf, QL],[Uf, 0]] -> [[T, 0],[Phi21, Phi22~]]
    double t[R][D], cc[R][D], e[R][D];
    {
      const double* M = publish<J0>(cx, 0, uf);
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
    const double* MT = publish<0>(cx, 1, t);
    {
      double dinv[R], inv[D];
#pragma unroll
      for (int s = 0; s < R; ++s) dinv[s] = fast_rcp(pick(t[s], cx.rc[s]));
      gather(cx, dinv, inv);
#pragma unroll
      for (int j = D - 1; j >= 0; --j) {
        double acc[R];
#pragma unroll
        for (int s = 0; s < R; ++s) acc[s] = e[s][j];
#pragma unroll
        for (int i = j + 1; i < D; ++i) {
          const double tv = MT[i * LDM + j];
#pragma unroll
          for (int s = 0; s < R; ++s) acc[s] = fma(-e[s][i], tv, acc[s]);
        }
#pragma unroll
        for (int s = 0; s < R; ++s) e[s][j] = acc[s] * inv[j];
      }
    }
    // ---- means: mp = F m ; g = m - E mp
    double mv[D], g[R];
    gather(cx, m, mv);
    mulF_vec(mv);
#pragma unroll
    for (int s = 0; s < R; ++s) {
      double acc = m[s];
#pragma unroll
      for (int i = 0; i < D; ++i) acc = fma(-e[s][i], mv[i], acc);
      g[s] = acc;
    }
    // ---- store the step's backward kernel (g | E | Phi22~): the noise factor stays UNtriangularised -- the smoother
    // only needs Phi22~ Phi22~^T and triangularises [E L | Phi22~] in one go (D - J0 - 1 pivots less per step here)
    {
      double* kp = kern + k * NE;
#pragma unroll
      for (int s = 0; s < R; ++s)
        if (cx.row[s] < D) kp[cx.row[s]] = g[s];
      store_rows_pm<0>(cx, kp + D, e);
      store_rows_pm<(J0 / PW) * PW>(cx, kp + D + D * D, uf);
    }
    // ---- measurement update
    double SL[d][d], y[d], zz[d];
    LinK lk;
    fetch_lin(cx, lin, k, lk);
    update(cx, lin, lk, MT, t, SL);
    H_times_vec(lin, lk, mv, y);
    solveSL(SL, y, zz);
#pragma unroll
    for (int s = 0; s < R; ++s) {
      double acc = pick(mv, cx.rc[s]);
#pragma unroll
      for (int a = 0; a < d; ++a) acc = fma(-t[s][a], zz[a], acc);
      m[s] = acc;
#pragma unroll
      for (int j = 0; j < D; ++j) uf[s][j] = (j < d) ? 0.0 : t[s][j];
    }
    // ---- innovation statistics (replicated)
    double q2 = 0.0, lg = 0.0;
#pragma unroll
    for (int a = 0; a < d; ++a) {
      q2 = fma(zz[a], zz[a], q2);
      lg += log(fabs(SL[a][a]));
    }
    st.nll += 0.5 * q2 + lg + 0.5 * d * LOG_2PI;
    st.s2 += q2;
    double wv[d], ww = 0.0;
#pragma unroll
    for (int a = d - 1; a >= 0; --a) {
      double acc = y[a];
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
                                              const double* __restrict__ state_in, double* __restrict__ kern,
                                              double* __restrict__ state_end, double* __restrict__ part,
                                              double* __restrict__ fmeans, double* __restrict__ fchols) {
    double m[R], uf[R][D];
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
    double le[R][D];
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
  static __device__ __forceinline__ void load_rows(const Ctx& cx, const double* __restrict__ base, double (&x)[R][D]) {
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const int rc = cx.rc[s];
#pragma unroll
      for (int j = 0; j < JS; ++j) x[s][j] = 0.0;
      if constexpr (PW == 2) {
        const double2* pe = reinterpret_cast<const double2*>(base) + rc;
#pragma unroll
        for (int j = JS / 2; j < D / 2; ++j) {
          const double2 a = pe[j * D];
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
  static __device__ __forceinline__ void store_rows_pm(const Ctx& cx, double* __restrict__ base,
                                                       const double (&x)[R][D]) {
#pragma unroll
    for (int s = 0; s < R; ++s)
      if (cx.row[s] < D) {
        const int r = cx.row[s];
        if constexpr (PW == 2) {
          double2* pe = reinterpret_cast<double2*>(base) + r;
#pragma unroll
          for (int j = JS / 2; j < D / 2; ++j) pe[j * D] = make_double2(x[s][2 * j], x[s][2 * j + 1]);
        } else {
#pragma unroll
          for (int j = JS; j < D; ++j) base[j * D + r] = x[s][j];
        }
      }
  }
  // L2 prefetch of one step's backward kernel (NE contiguous doubles) and the previous mean of that row, spread over
  // the group's lanes in 128-byte strides
  static __device__ __forceinline__ void prefetch_step(const Ctx& cx, const double* __restrict__ kp,
                                                       const double* __restrict__ mrow) {
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
  static __device__ __forceinline__ double emit(const Ctx& cx, long t, const double (&m)[R], const double (&l)[R][D],
                                                double cscale, const double (&old)[R], double* __restrict__ means,
                                                double* __restrict__ chols) {
    double bad = 0.0;
#pragma unroll
    for (int s = 0; s < R; ++s)
      if (cx.row[s] < D) {
        const int r = cx.row[s];
        const bool close = fabs(old[s] - m[s]) <= (1e-8 + 1e-13 * fabs(m[s]));
        bad += close ? 0.0 : 1.0;
        means[t * D + r] = m[s];
        if (chols) {
          double rowv[D];
#pragma unroll
          for (int j = 0; j < D; ++j) rowv[j] = (j <= r) ? cscale * l[s][j] : 0.0;
          store_row(chols + (t * D + r) * D, rowv);
        }
      }
    return bad;
  }
  template <int JS>
  static __device__ __forceinline__ void smooth_step(Ctx& cx, long k, bool has_prev, bool emit_t0,
                                                     const double* qLinvdiag, const double* qL,
                                                     const double* __restrict__ kern, double cscale,
                                                     double* __restrict__ means, double* __restrict__ chols,
                                                     double (&m)[R], double (&l)[R][D], double& obj, double& bad) {
    double g[R], e[R][D], ph[R][D], old[R];
    {
      const double* kp = kern + k * NE;
#pragma unroll
      for (int s = 0; s < R; ++s) {
        g[s] = kp[cx.rc[s]];
        old[s] = (k > 0 || emit_t0) ? means[k * D + cx.rc[s]] : 0.0;
      }
      load_rows<0>(cx, kp + D, e);
      load_rows<JS>(cx, kp + D + D * D, ph);
    }
    if (has_prev) prefetch_step(cx, kern + (k - 1) * NE, means + (k - 1) * D);
    const double* ML = publish<0>(cx, 0, l);
    double mv[D];
    gather(cx, m, mv);
    double mn[R], cd[R][D];
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
        const double lv = ML[i * LDM + j];
#pragma unroll
        for (int s = 0; s < R; ++s) cd[s][j] = fma(e[s][i], lv, cd[s][j]);
      }
    }
    tria2_step<0, JS>(cx, cd, ph);
    // objective increment |QL^{-1}(m_k - F m_{k+1})|^2 (replicated)
    double fm[D], rr[D], dr[R];
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
        double acc = rr[b * Q1 + i];
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
                                                const double* qLinvdiag, const double* qL,
                                                const double* __restrict__ seed, const double* __restrict__ kern,
                                                double cscale, double* __restrict__ means,
                                                double* __restrict__ chols, double* __restrict__ part) {
    double m[R], l[R][D], old[R];
#pragma unroll
    for (int s = 0; s < R; ++s) {
      const bool ok = cx.row[s] < D;
      m[s] = ok ? seed[cx.rc[s]] : 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) l[s][j] = (ok && j <= cx.rc[s]) ? seed[D + cx.rc[s] * D + j] : 0.0;
    }
    double obj = 0.0, bad = 0.0;
    if (last) {
#pragma unroll
      for (int s = 0; s < R; ++s) old[s] = means[k1 * D + cx.rc[s]];
      bad += emit(cx, k1, m, l, cscale, old, means, chols);
    }
    // Global loads of a step are issued at its top and hit L2: the whole kernel of step k-1 (and the previous mean of
    // row k-1) is prefetched into L2 while step k computes.  (Holding the next kernel in registers spilled; mixing
    // DRAM-latency loads for step k-1 with L2 hits for step k made every consumer wait for the slowest load, because
    // the few load scoreboards are shared -- ncu: long_scoreboard on the first shuffle of Dk, 33 % of the samples.)
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

}  // namespace pof
