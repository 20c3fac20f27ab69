// Leaf level, performance path: G lanes of a warp cooperate on ONE time-chunk, lane r owning ROW r of every
// D-row matrix of the recursion (G = smallest power of two >= D, 32/G chunks per warp).
//
// Why this shape (B200): the recursions are chains of small Householder triangularisations applied from the right.
// With one matrix row per lane every reflection is embarrassingly row-parallel -- the pivot row is published once
// through shared memory (one STS by the owner, broadcast LDS by the group) and each lane then updates only its own
// registers with FP64 FMAs: no cross-lane reductions, no shuffles in the inner loops, all loops unrolled over
// compile-time column indices so the rows live in registers without spilling (a thread-per-chunk layout needs
// ~150-300 live doubles per thread at D = 8 and spills; it also takes nvcc half an hour to unroll).
// Global traffic is row-contiguous per lane, i.e. coalesced 8*D-byte segments per group.
//
// The math is the same as the thread-private reference implementation in pof_leaf.cuh (validated against the CPU
// oracle in tests/hostsim); see that file and pof_pipeline.cuh for the reference formulas and citations.
#pragma once
#include <cuda_runtime.h>

#include "pof_small.cuh"

namespace pof {

constexpr int pow2ceil(int x) { return x <= 1 ? 1 : (x <= 2 ? 2 : (x <= 4 ? 4 : (x <= 8 ? 8 : (x <= 16 ? 16 : 32)))); }

template <int d, int q>
struct Lane {
  static constexpr int Q1 = q + 1;
  static constexpr int D = d * Q1;
  static constexpr int G = pow2ceil(D);
  static constexpr int GPW = 32 / G;  // chunks (groups) per warp
  static constexpr int NE = D + 2 * D * D;
  static constexpr double LOG_2PI = 1.8378770664093454835606594728112;
  // per-group shared memory (doubles): two broadcast slots + two row-exchange matrices (+ padding so that the groups
  // of a warp start 2 doubles (mod 16) apart: conflict-free broadcast reads across groups)
  static constexpr int LDM = D + 1;
  static constexpr int BC = ((D + 2 + 1) / 2) * 2;
  static constexpr int VEC = ((D + 1) / 2) * 2;
  static constexpr int RAW = 2 * BC + 2 * D * LDM + 2 * VEC;
  static constexpr int SM_GROUP = RAW + ((2 - (RAW % 16)) + 16) % 16;

  struct Ctx {
    int r;          // row owned by this lane (lane index within the group); rows >= D are idle
    int rb, blk0;   // row within its block, first row of the block
    unsigned mask;  // lanes of this group
    double* bc;     // broadcast slots
    double* mat;    // row-exchange matrices
    double* vec;    // gather vector
    int flip, vflip;
    double cf[Q1];  // Pascal coefficients of this lane's row of F: (F x)_r = sum_i cf[i] x[blk0 + i]
    double tq[D];   // this lane's row of QL
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
  };

  static __device__ __forceinline__ void init_ctx(Ctx& c, double* sm_group, const double* qL) {
    const int lane = threadIdx.x & 31;
    c.r = lane % G;
    c.mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
    c.bc = sm_group;
    c.mat = sm_group + 2 * BC;
    c.vec = c.mat + 2 * D * LDM;
    c.flip = 0;
    c.vflip = 0;
    const int rr = (c.r < D) ? c.r : 0;
    c.rb = rr % Q1;
    c.blk0 = rr - c.rb;
#pragma unroll
    for (int i = 0; i < Q1; ++i) {
      double v = 0.0;
#pragma unroll
      for (int b = 0; b < Q1; ++b)
        if (c.rb == b && i >= b) v = binom(q - b, i - b);
      c.cf[i] = v;
    }
#pragma unroll
    for (int j = 0; j < D; ++j) {
      double v = 0.0;
#pragma unroll
      for (int b = 0; b < Q1; ++b)
        if (c.rb == b && (j / Q1) * Q1 == c.blk0 && (j % Q1) <= b) v = qL[b * Q1 + (j % Q1)];
      c.tq[j] = (c.r < D) ? v : 0.0;
    }
  }

  // ---- communication primitives (group scope) -----------------------------------------------------------------
  // lane `src` publishes n doubles, every lane of the group reads them
  template <int n>
  static __device__ __forceinline__ void bcast(Ctx& c, const double (&x)[n], int src, double (&out)[n]) {
#ifndef POF_BCAST_SMEM
    // register-to-register broadcast (default; measured 2 % faster than the shared-memory slot on B200: the same
    // crossbar bandwidth, but no store + barrier on the critical path)
    const int srclane = ((threadIdx.x & 31) / G) * G + src;
#pragma unroll
    for (int j = 0; j < n; ++j) out[j] = __shfl_sync(c.mask, x[j], srclane);
    return;
#endif
    double2* slot = reinterpret_cast<double2*>(c.bc + c.flip * BC);
    c.flip ^= 1;
    if (c.r == src) {
#pragma unroll
      for (int j = 0; j + 1 < n; j += 2) slot[j / 2] = make_double2(x[j], x[j + 1]);
      if (n % 2) slot[n / 2] = make_double2(x[n - 1], 0.0);
    }
    c.sync();
#pragma unroll
    for (int j = 0; j + 1 < n; j += 2) {
      const double2 v = slot[j / 2];
      out[j] = v.x;
      out[j + 1] = v.y;
    }
    if (n % 2) out[n - 1] = slot[n / 2].x;
  }
  // out[j] = x of lane j
  static __device__ __forceinline__ void allgather(Ctx& c, double x, double (&out)[D]) {
    double* v = c.vec + c.vflip * VEC;
    c.vflip ^= 1;
    if (c.r < D) v[c.r] = x;
    c.sync();
#pragma unroll
    for (int j = 0; j < D; ++j) out[j] = v[j];
  }
  // publish this lane's row into exchange matrix `which` (row-major, leading dimension LDM)
  static __device__ __forceinline__ double* publish(Ctx& c, int which, const double (&x)[D]) {
    double* M = c.mat + which * D * LDM;
    c.sync();
    if (c.r < D) {
#pragma unroll
      for (int j = 0; j < D; ++j) M[c.r * LDM + j] = x[j];
    }
    c.sync();
    return M;
  }
  // y = row r of (F X) given the published rows of X
  static __device__ __forceinline__ void mulF_row(const Ctx& c, const double* M, double (&y)[D]) {
#pragma unroll
    for (int j = 0; j < D; ++j) y[j] = 0.0;
#pragma unroll
    for (int i = 0; i < Q1; ++i) {
      const double* row = M + (c.blk0 + i) * LDM;
#pragma unroll
      for (int j = 0; j < D; ++j) y[j] = fma(c.cf[i], row[j], y[j]);
    }
  }
  // (F v)_r from the gathered vector v
  static __device__ __forceinline__ double mulF_at(const Ctx& c, const double (&v)[D]) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < Q1; ++i) {
      double x = 0.0;
#pragma unroll
      for (int b = 0; b < d; ++b)
        if (c.blk0 == b * Q1) x = v[b * Q1 + i];
      s = fma(c.cf[i], x, s);
    }
    return s;
  }
  // full F v (redundant in every lane)
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

  // reciprocal and reciprocal square root: hardware seed (2^-22) + two Newton steps (~1 ulp), branch free
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

  // Householder parameters for the row (alpha, x[0..n-1]):  H = I - tp * v v^T,  v = (s, x),  H (alpha,x)^T = (beta,0)
  struct HH {
    double s, tp, beta;
    bool nz;
  };
  template <int n>
  static __device__ __forceinline__ HH house(double alpha, const double* x) {
    double sigma = 0.0;
#pragma unroll
    for (int j = 0; j < n; ++j) sigma = fma(x[j], x[j], sigma);
    HH h;
    const double nrm2 = fma(alpha, alpha, sigma);
    // rsqrt.approx.ftz flushes subnormal inputs to zero (-> inf -> NaN in the Newton step): a row whose squared norm
    // is below 2^-1000 is treated as already reduced (identity reflector), as a zero tail is
    h.nz = sigma > 0.0 && nrm2 > 0x1p-1000;
    const double rn = fast_rsqrt(nrm2);  // 1 / norm
    const double nrm = nrm2 * rn;
    const double beta = (alpha >= 0.0) ? -nrm : nrm;
    const double s = alpha - beta;
    h.beta = h.nz ? beta : alpha;
    h.s = h.nz ? s : 0.0;
    h.tp = h.nz ? rn * fast_rcp(fabs(s)) : 0.0;  // 1 / (norm |s|)
    return h;
  }

  // Triangular-pentagonal right-QR, rows over lanes.  t: row r of T (lower triangular, D x D), c: row r of C (D x K).
  // Passenger rows (pt: D entries, pc: K entries) see the same reflections.
  template <int K, bool PASS>
  static __device__ __forceinline__ void tpqrt(Ctx& cx, double (&t)[D], double (&c)[K], double* pt, double* pc) {
#pragma unroll
    for (int i = 0; i < D; ++i) {
      double mine[K + 1], piv[K + 1];
      mine[0] = t[i];
#pragma unroll
      for (int j = 0; j < K; ++j) mine[1 + j] = c[j];
      bcast<K + 1>(cx, mine, i, piv);
      const HH h = house<K>(piv[0], piv + 1);
      double w = h.s * t[i];
#pragma unroll
      for (int j = 0; j < K; ++j) w = fma(c[j], piv[1 + j], w);
      // rows above the pivot are untouched (w = 0); the pivot row itself maps to (beta, ~0): its C entries are dead
      w = (cx.r >= i) ? w * h.tp : 0.0;
      t[i] = (cx.r == i) ? h.beta : fma(-w, h.s, t[i]);
#pragma unroll
      for (int j = 0; j < K; ++j) c[j] = fma(-w, piv[1 + j], c[j]);
      if (PASS) {
        double u = h.s * pt[i];
#pragma unroll
        for (int j = 0; j < K; ++j) u = fma(pc[j], piv[1 + j], u);
        u *= h.tp;
        pt[i] = fma(-u, h.s, pt[i]);
#pragma unroll
        for (int j = 0; j < K; ++j) pc[j] = fma(-u, piv[1 + j], pc[j]);
      }
    }
  }

  // plain right-Householder lower-triangularisation of a D x D matrix, rows over lanes
  template <int I>
  static __device__ __forceinline__ void tria_step(Ctx& cx, double (&x)[D]) {
    if constexpr (I + 1 < D) {
      constexpr int n = D - I;
      double mine[n], piv[n];
#pragma unroll
      for (int j = I; j < D; ++j) mine[j - I] = x[j];
      bcast<n>(cx, mine, I, piv);
      const HH h = house<n - 1>(piv[0], piv + 1);
      double w = h.s * x[I];
#pragma unroll
      for (int j = I + 1; j < D; ++j) w = fma(x[j], piv[j - I], w);
      // entries right of the diagonal in the pivot row become ~0 (roundoff); only the lower triangle is ever read
      w = (cx.r >= I) ? w * h.tp : 0.0;
      x[I] = (cx.r == I) ? h.beta : fma(-w, h.s, x[I]);
#pragma unroll
      for (int j = I + 1; j < D; ++j) x[j] = fma(-w, piv[j - I], x[j]);
      tria_step<I + 1>(cx, x);
    }
  }
  static __device__ __forceinline__ void tria_rows(Ctx& cx, double (&x)[D]) { tria_step<0>(cx, x); }

  // select x[idx] for a runtime idx without dynamic register indexing
  static __device__ __forceinline__ double pick(const double (&x)[D], int idx) {
    double v = 0.0;
#pragma unroll
    for (int i = 0; i < D; ++i)
      if (i == idx) v = x[i];
    return v;
  }

  template <int a>
  static __device__ __forceinline__ void update_pivot(double (&t)[D], double (&W)[d][D]) {
    if constexpr (a < d) {
      const HH h = house<D - a - 1>(W[a][a], &W[a][a + 1]);
      // own row of T
      double w = h.s * t[a];
#pragma unroll
      for (int j = a + 1; j < D; ++j) w = fma(t[j], W[a][j], w);
      w *= h.tp;
      t[a] = fma(-w, h.s, t[a]);
#pragma unroll
      for (int j = a + 1; j < D; ++j) t[j] = fma(-w, W[a][j], t[j]);
      // remaining pivot rows (replicated in every lane)
#pragma unroll
      for (int a2 = a + 1; a2 < d; ++a2) {
        double u = h.s * W[a2][a];
#pragma unroll
        for (int j = a + 1; j < D; ++j) u = fma(W[a2][j], W[a][j], u);
        u *= h.tp;
        W[a2][a] = fma(-u, h.s, W[a2][a]);
#pragma unroll
        for (int j = a + 1; j < D; ++j) W[a2][j] = fma(-u, W[a][j], W[a2][j]);
      }
      W[a][a] = h.beta;
      update_pivot<a + 1>(t, W);
    }
  }

  // Measurement update on the predicted factor.  In: t = row r of T (lower triangular, full D entries with zeros
  // above the diagonal), H (d x D, replicated).  Out: SL (replicated), kbar = row r of Kbar, t = row r of the
  // posterior factor (first d entries are the Kbar entries -> caller zeroes them).
  static __device__ __forceinline__ void update(Ctx& cx, double (&t)[D], const double (&H)[d][D], double (&SL)[d][d]) {
    // W[a][r] = sum_i H[a][i] T[i][r]: column r of T from the published rows
    const double* M = publish(cx, 0, t);
    double wcol[d];
#pragma unroll
    for (int a = 0; a < d; ++a) wcol[a] = 0.0;
    const int rc = (cx.r < D) ? cx.r : 0;
#pragma unroll
    for (int i = 0; i < D; ++i) {
      const double tv = M[i * LDM + rc];
#pragma unroll
      for (int a = 0; a < d; ++a) wcol[a] = fma(H[a][i], tv, wcol[a]);
    }
    double W[d][D];
#pragma unroll
    for (int a = 0; a < d; ++a) allgather(cx, wcol[a], W[a]);
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
  // linearised observation model of step k: dense (H, c) from memory, or rebuilt from the compact form [J_f | c]:
  // H = E1 - J_f E0 with E0 = s0 e_0^T, E1 = s1 e_1^T per block (reference pof/convenience.py:26-28)
  struct Lin {
    const double* __restrict__ H;
    const double* __restrict__ c;
    const double* __restrict__ Jc;
    double s0, s1;
  };
  static __device__ __forceinline__ void load_Hc(const Lin& L, long k, double (&Hk)[d][D], double (&ck)[d]) {
    if (L.Jc) {
      const double* p = L.Jc + k * (d * d + d);
#pragma unroll
      for (int a = 0; a < d; ++a) {
        ck[a] = __ldg(p + d * d + a);
#pragma unroll
        for (int j = 0; j < D; ++j) Hk[a][j] = 0.0;
#pragma unroll
        for (int b = 0; b < d; ++b) Hk[a][b * Q1] = -__ldg(p + a * d + b) * L.s0;
        Hk[a][a * Q1 + 1] += L.s1;
      }
    } else {
#pragma unroll
      for (int a = 0; a < d; ++a) {
        ck[a] = __ldg(L.c + k * d + a);
#pragma unroll
        for (int j = 0; j < D; ++j) Hk[a][j] = __ldg(L.H + (k * d + a) * D + j);
      }
    }
  }

  // ================================================================== filter phase 1: chunk -> filtering element
  static __device__ __forceinline__ void fold(Ctx& cx, long k0, long k1, const Lin& lin, double* __restrict__ agg,
                                              double* __restrict__ aggm) {
    const int r = cx.r;
    double a[D], uf[D], z[D], b = 0.0, eta = 0.0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      a[j] = (j == r) ? 1.0 : 0.0;
      uf[j] = 0.0;
      z[j] = 0.0;
    }
    for (long k = k0; k < k1; ++k) {
      double Hk[d][D], ck[d];
      load_Hc(lin, k, Hk, ck);
      // predict
      double t[D], cc[D];
      {
        const double* M = publish(cx, 0, a);
        mulF_row(cx, M, a);
        const double* M2 = publish(cx, 1, uf);
        mulF_row(cx, M2, cc);
        double bv[D];
        allgather(cx, b, bv);
        b = mulF_at(cx, bv);
      }
#pragma unroll
      for (int j = 0; j < D; ++j) t[j] = cx.tq[j];
      tpqrt<D, false>(cx, t, cc, nullptr, nullptr);
      // the element before the chunk's last measurement update (for the chunk-level smoothing element)
      if (aggm && k == k1 - 1 && r < D) {
        const int DD = D * D;
        aggm[DD + r] = b;
        aggm[2 * DD + D + r] = eta;
#pragma unroll
        for (int j = 0; j < D; ++j) {
          aggm[r * D + j] = a[j];
          aggm[DD + D + r * D + j] = (j <= r) ? t[j] : 0.0;
          aggm[2 * DD + 2 * D + r * D + j] = (j <= r) ? z[j] : 0.0;
        }
      }
      // update
      double SL[d][d];
      update(cx, t, Hk, SL);
      // HA[:, r], H b + c
      double bv[D];
      allgather(cx, b, bv);
      const double* MA = publish(cx, 1, a);
      double g[d], rv[d], zz[d];
      const int rc = (r < D) ? r : 0;
#pragma unroll
      for (int e = 0; e < d; ++e) {
        g[e] = 0.0;
        rv[e] = ck[e];
      }
#pragma unroll
      for (int i = 0; i < D; ++i) {
        const double av = MA[i * LDM + rc];
#pragma unroll
        for (int e = 0; e < d; ++e) {
          g[e] = fma(Hk[e][i], av, g[e]);
          rv[e] = fma(Hk[e][i], bv[i], rv[e]);
        }
      }
      solveSL(SL, rv, zz);
      {
        double gg[d];
        solveSL(SL, g, gg);
#pragma unroll
        for (int e = 0; e < d; ++e) g[e] = gg[e];
      }
      // A <- A - Kbar G ; b <- b - Kbar z ; eta <- eta - G^T z
      double Gf[d][D];
#pragma unroll
      for (int e = 0; e < d; ++e) allgather(cx, g[e], Gf[e]);
#pragma unroll
      for (int e = 0; e < d; ++e) {
        const double kb = t[e];
        b = fma(-kb, zz[e], b);
        eta = fma(-g[e], zz[e], eta);
#pragma unroll
        for (int j = 0; j < D; ++j) a[j] = fma(-kb, Gf[e][j], a[j]);
      }
#pragma unroll
      for (int j = 0; j < D; ++j) uf[j] = (j < d) ? 0.0 : t[j];
      // Z <- tria([Z, G^T])
      tpqrt<d, false>(cx, z, g, nullptr, nullptr);
    }
    if (r < D) {
      const int DD = D * D;
      agg[DD + r] = b;
      agg[2 * DD + D + r] = eta;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        agg[r * D + j] = a[j];
        agg[DD + D + r * D + j] = uf[j];
        agg[2 * DD + 2 * D + r * D + j] = (j <= r) ? z[j] : 0.0;
      }
    }
  }

  // ================================================================== filter phase 3: seeded square-root KF
  template <bool COMPOSE>
  static __device__ __forceinline__ void scan(Ctx& cx, long k0, long k1, const Lin& lin,
                                              const double* __restrict__ state_in,
                                              double* __restrict__ kern, double* __restrict__ sagg,
                                              double* __restrict__ state_end, double* __restrict__ part,
                                              double* __restrict__ fmeans, double* __restrict__ fchols) {
    const int r = cx.r;
    const int rc = (r < D) ? r : 0;
    double m = state_in[rc], uf[D];
#pragma unroll
    for (int j = 0; j < D; ++j) uf[j] = state_in[D + rc * D + j];
    if (r >= D) {
      m = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) uf[j] = 0.0;
    }
    double ga = 0.0, ea[D], da[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      ea[j] = 0.0;
      da[j] = 0.0;
    }
    double nll = 0.0, s1 = 0.0, s2 = 0.0;
    for (long k = k0; k < k1; ++k) {
      double Hk[d][D], ck[d];
      load_Hc(lin, k, Hk, ck);
      // ---- predict + backward kernel: [[F Uf, QL],[Uf, 0]] -> [[T, 0],[Phi21, Phi22~]]
      double t[D], cc[D], e[D];
      {
        const double* M = publish(cx, 0, uf);
        mulF_row(cx, M, cc);
      }
#pragma unroll
      for (int j = 0; j < D; ++j) {
        t[j] = cx.tq[j];
        e[j] = 0.0;
      }
      tpqrt<D, true>(cx, t, cc, e, uf);
      // ---- E row: e <- e T^{-1}
      {
        const double* M = publish(cx, 1, t);
        double inv[D];
        allgather(cx, fast_rcp(pick(t, rc)), inv);
#pragma unroll
        for (int j = D - 1; j >= 0; --j) {
          double s = e[j];
#pragma unroll
          for (int i = j + 1; i < D; ++i) s = fma(-e[i], M[i * LDM + j], s);
          e[j] = s * inv[j];
        }
      }
      // ---- means: mp = F m ; g_r = m_r - E[r,:] mp
      double mv[D];
      allgather(cx, m, mv);
      mulF_vec(mv);
      double g = m;
#pragma unroll
      for (int i = 0; i < D; ++i) g = fma(-e[i], mv[i], g);
      // ---- Dk = tria(Phi22~)
      tria_rows(cx, uf);
      // ---- store the step's backward kernel (time-major, row-contiguous per lane)
      if (r < D) {
        double* kp = kern + k * NE;
        kp[r] = g;
        store_row(kp + D + r * D, e);
        store_row(kp + D + D * D + r * D, uf);
      }
      // ---- compose the chunk's smoothing element: acc = acc o kernel_k  (skipped when the chunk-level element is
      //      derived from the filtering element instead, pof_treelane.cuh chunk_kernel)
      if (!COMPOSE) {
      } else if (k == k0) {
        ga = g;
#pragma unroll
        for (int j = 0; j < D; ++j) {
          ea[j] = e[j];
          da[j] = uf[j];
        }
      } else {
        double gv[D];
        allgather(cx, g, gv);
        const double* ME = publish(cx, 0, e);
        const double* MD = publish(cx, 1, uf);
        double ne[D], cd[D];
#pragma unroll
        for (int j = 0; j < D; ++j) {
          ne[j] = 0.0;
          cd[j] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < D; ++i) {
          ga = fma(ea[i], gv[i], ga);
#pragma unroll
          for (int j = 0; j < D; ++j) {
            ne[j] = fma(ea[i], ME[i * LDM + j], ne[j]);
            if (j <= i) cd[j] = fma(ea[i], MD[i * LDM + j], cd[j]);
          }
        }
#pragma unroll
        for (int j = 0; j < D; ++j) ea[j] = ne[j];
        tpqrt<D, false>(cx, da, cd, nullptr, nullptr);
      }
      // ---- measurement update
      double SL[d][d], y[d], zz[d];
      update(cx, t, Hk, SL);
#pragma unroll
      for (int a = 0; a < d; ++a) {
        double s = ck[a];
#pragma unroll
        for (int i = 0; i < D; ++i) s = fma(Hk[a][i], mv[i], s);
        y[a] = s;
      }
      solveSL(SL, y, zz);
      m = pick(mv, rc);
#pragma unroll
      for (int a = 0; a < d; ++a) m = fma(-t[a], zz[a], m);
#pragma unroll
      for (int j = 0; j < D; ++j) uf[j] = (j < d) ? 0.0 : t[j];
      // ---- innovation statistics (replicated; lane 0 reports)
      double q2 = 0.0, lg = 0.0;
#pragma unroll
      for (int a = 0; a < d; ++a) {
        q2 = fma(zz[a], zz[a], q2);
        lg += log(fabs(SL[a][a]));
      }
      nll += 0.5 * q2 + lg + 0.5 * d * LOG_2PI;
      s2 += q2;
      double wv[d], ww = 0.0;
#pragma unroll
      for (int a = d - 1; a >= 0; --a) {
        double s = y[a];
#pragma unroll
        for (int e2 = a + 1; e2 < d; ++e2) s = fma(-SL[e2][a], wv[e2], s);
        wv[a] = s * fast_rcp(SL[a][a]);
        ww = fma(wv[a], wv[a], ww);
      }
      s1 += ww;
      if (fmeans && r < D) {
        fmeans[(k + 1) * D + r] = m;
#pragma unroll
        for (int j = 0; j < D; ++j) fchols[((k + 1) * D + r) * D + j] = uf[j];
      }
    }
    // chunk outputs: smoothing element, filtered end state (triangular factor), partial sums
    tria_rows(cx, uf);
    if (r < D) {
      state_end[r] = m;
#pragma unroll
      for (int j = 0; j < D; ++j) state_end[D + r * D + j] = (j <= r) ? uf[j] : 0.0;
      if (COMPOSE) {
        sagg[r] = ga;
#pragma unroll
        for (int j = 0; j < D; ++j) {
          sagg[D + r * D + j] = ea[j];
          sagg[D + D * D + r * D + j] = (j <= r) ? da[j] : 0.0;
        }
      }
    }
    if (r == 0) {
      part[0] = nll;
      part[1] = s1;
      part[2] = s2;
    }
  }

  // ================================================================== smoother phase 3: seeded square-root RTS
  static __device__ __forceinline__ double emit(int r, long t, double m, const double (&l)[D], double cscale,
                                                double old, double* __restrict__ means,
                                                double* __restrict__ chols) {
    if (r >= D) return 0.0;
    const bool close = fabs(old - m) <= (1e-8 + 1e-13 * fabs(m));
    means[t * D + r] = m;
    if (chols) {
      double row[D];
#pragma unroll
      for (int j = 0; j < D; ++j) row[j] = (j <= r) ? cscale * l[j] : 0.0;
      store_row(chols + (t * D + r) * D, row);
    }
    return close ? 0.0 : 1.0;
  }
  // one matrix row to global memory (128-bit stores when the row length is even)
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
  static __device__ __forceinline__ void load_kernel(int r, const double* __restrict__ kp, double& g, double (&e)[D],
                                                     double (&dk)[D]) {
    const int rc = (r < D) ? r : 0;
    g = kp[rc];
    const double2* pe = reinterpret_cast<const double2*>(kp + D + rc * D);
    const double2* pd = reinterpret_cast<const double2*>(kp + D + D * D + rc * D);
    if constexpr (D % 2 == 0) {
#pragma unroll
      for (int j = 0; j < D / 2; ++j) {
        const double2 a = pe[j], b = pd[j];
        e[2 * j] = a.x;
        e[2 * j + 1] = a.y;
        dk[2 * j] = b.x;
        dk[2 * j + 1] = b.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < D; ++j) {
        e[j] = kp[D + rc * D + j];
        dk[j] = kp[D + D * D + rc * D + j];
      }
    }
  }
  static __device__ __forceinline__ void smooth(Ctx& cx, long k0, long k1, bool last, bool emit_t0,
                                                const double* qLinvdiag, const double* qL,
                                                const double* __restrict__ seed, const double* __restrict__ kern,
                                                double cscale, double* __restrict__ means,
                                                double* __restrict__ chols, double* __restrict__ part) {
    const int r = cx.r;
    const int rc = (r < D) ? r : 0;
    double m = seed[rc], l[D];
#pragma unroll
    for (int j = 0; j < D; ++j) l[j] = (j <= rc) ? seed[D + rc * D + j] : 0.0;
    if (r >= D) {
      m = 0.0;
#pragma unroll
      for (int j = 0; j < D; ++j) l[j] = 0.0;
    }
    double obj = 0.0, bad = 0.0;
    if (last) bad += emit(r, k1, m, l, cscale, means[k1 * D + rc], means, chols);
    // software pipeline: the backward kernel and the previous mean of step k-1 are in flight while step k is
    // processed (no dependent global load inside a step)
    double gn = 0.0, en[D], dkn[D];
    load_kernel(r, kern + (k1 - 1) * NE, gn, en, dkn);
    const bool skip0 = !emit_t0;  // row t = 0 belongs to the previous shard
    double oldn = (k1 - 1 > 0 || !skip0) ? means[(k1 - 1) * D + rc] : 0.0;
    for (long k = k1 - 1; k >= k0; --k) {
      double g = gn, e[D], dk[D];
      const double old = oldn;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        e[j] = en[j];
        dk[j] = dkn[j];
      }
      if (k > k0) {
        load_kernel(r, kern + (k - 1) * NE, gn, en, dkn);
        oldn = (k - 1 > 0 || !skip0) ? means[(k - 1) * D + rc] : 0.0;
      }
      const double* ML = publish(cx, 0, l);
      double mv[D];
      allgather(cx, m, mv);
      double mn = g, cd[D];
#pragma unroll
      for (int j = 0; j < D; ++j) cd[j] = 0.0;
#pragma unroll
      for (int i = 0; i < D; ++i) {
        mn = fma(e[i], mv[i], mn);
#pragma unroll
        for (int j = 0; j <= i; ++j) cd[j] = fma(e[i], ML[i * LDM + j], cd[j]);
      }
      tpqrt<D, false>(cx, dk, cd, nullptr, nullptr);
      // objective increment |QL^{-1}(m_k - F m_{k+1})|^2, replicated
      double rr[D];
      allgather(cx, mn - mulF_at(cx, mv), rr);
#pragma unroll
      for (int b = 0; b < d; ++b) {
#pragma unroll
        for (int i = 0; i < Q1; ++i) {
          double s = rr[b * Q1 + i];
#pragma unroll
          for (int j = 0; j < i; ++j) s = fma(-qL[i * Q1 + j], rr[b * Q1 + j], s);
          s *= qLinvdiag[i];
          rr[b * Q1 + i] = s;
          obj = fma(s, s, obj);
        }
      }
      m = mn;
#pragma unroll
      for (int j = 0; j < D; ++j) l[j] = (j <= rc) ? dk[j] : 0.0;
      if (k > 0 || emit_t0) bad += emit(r, k, m, l, cscale, old, means, chols);
    }
    // not-close counts are per lane: sum over the group
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) bad += __shfl_xor_sync(cx.mask, bad, o, G);
    if (r == 0) {
      part[0] = obj;
      part[1] = bad;
    }
  }
};

}  // namespace pof
