// Carry level, performance path: the reference's two associative operators with REGISTER-resident rows.
//
//   filter operator  (pof/parallel_filtsmooth/filter.py:117-142): G2 = pow2ceil(2D) lanes per combine, lane r owns
//                    row r of the 2D x 2D array Xi = [[U1^T Z2, I],[Z2, 0]] during its triangularisation; afterwards
//                    lanes 0..D-1 triangularise [A2 Y | U2] (-> U) while lanes D..2D-1 triangularise
//                    [A1^T Xi22 | Z1] (-> Z) in the same pivot loop.
//   smoothing operator (smoother.py:53-63): GS = pow2ceil(D) lanes per combine.
//
// Same algebra as pof_coop.cuh (Y = U1 Xi11^{-T}, G = I - Y Xi21^T, ...), same packed element layouts; the generic
// shared-memory version there remains the fallback for 2D > 32.  Small GEMMs read the second operand's rows from a
// per-combine shared-memory staging area (published row-wise by the owning lanes).
#pragma once
#include <cuda_runtime.h>

#include "pof_real.cuh"
#include "pof_small.cuh"

namespace POF_NS {

template <int D>
struct TreeLane {
  static constexpr int pow2c(int x) { return x <= 2 ? 2 : (x <= 4 ? 4 : (x <= 8 ? 8 : (x <= 16 ? 16 : 32))); }
  static constexpr int W2 = 2 * D;
  static constexpr int G2 = pow2c(W2);     // lanes per filtering combine
  static constexpr int GS = pow2c(D);      // lanes per smoothing combine
  static constexpr int LDM = D + 1;
  static constexpr int MAT = D * LDM;
  static constexpr int BCW = ((W2 + 2 + 1) / 2) * 2;  // broadcast slot (doubles)
  // staging: 9 matrices + 2x2 broadcast slots + 2 gather vectors
  static constexpr int NMAT = 9;
  static constexpr int VEC = ((W2 + 1) / 2) * 2;
  static constexpr int RAW = NMAT * MAT + 4 * BCW + 2 * VEC;
  static constexpr int SM_COMBINE = RAW + ((2 - (RAW % 16)) + 16) % 16;
  enum { mU1 = 0, mZ2, mA1, mX11, mX21, mX22, mY, mG, mP };
  // the smoothing operator stages only two matrices (D1, E1): a quarter of the shared memory, so that the smoother's
  // sweeps fit next to the two resident CTAs of the filter scan they run concurrently with (pof_api.cu, stage B)
  static constexpr int NMAT_S = 2;
  static constexpr int RAW_S = NMAT_S * MAT + 4 * BCW + 2 * VEC;
  static constexpr int SM_COMBINE_S = RAW_S + ((2 - (RAW_S % 16)) + 16) % 16;
  enum { sD1 = 0, sE1 = 1 };

  struct Ctx {
    int r;  // lane within the combine's group
    unsigned mask;
    real* sm;
    int nmat;  // matrices staged ahead of the broadcast slots (NMAT, or NMAT_S for the smoothing operator)
    int flip, vflip;
    __device__ __forceinline__ real* mat(int which) const { return sm + which * MAT; }
    __device__ __forceinline__ void sync() const { __syncwarp(mask); }
  };
  template <int G>
  static __device__ __forceinline__ void init(Ctx& c, real* sm, int nmat = NMAT) {
    const int lane = threadIdx.x & 31;
    c.nmat = nmat;
    c.r = lane % G;
    c.mask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << ((lane / G) * G));
    c.sm = sm;
    c.flip = 0;
    c.vflip = 0;
  }

  static __device__ __forceinline__ real fast_rcp(real x) { return POF_NS::fast_rcp(x); }
  static __device__ __forceinline__ real fast_rsqrt(real x) { return POF_NS::fast_rsqrt(x); }
  // multi-accumulator dot product (the combines are latency chains: split every long FMA chain)
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
  template <int n>
  static __device__ __forceinline__ HH house(real alpha, const real* x) {
    const real sigma = dotn<n>(x, x);
    const real nrm2 = fma(alpha, alpha, sigma);
    const bool nz = sigma > real(0) && nrm2 > tiny_norm2();  // see pof_real.cuh: subnormal norms would turn into NaN
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

  // lane `src` publishes n doubles into broadcast slot `slot_id` (0/1: two independent streams), all lanes read
  template <int n>
  static __device__ __forceinline__ void bcast(Ctx& c, int slot_id, const real (&x)[n], int src, real (&out)[n]) {
    real2* slot = reinterpret_cast<real2*>(c.sm + c.nmat * MAT + (2 * slot_id + c.flip) * BCW);
    if (c.r == src) {
#pragma unroll
      for (int j = 0; j + 1 < n; j += 2) slot[j / 2] = make_real2(x[j], x[j + 1]);
      if (n % 2) slot[n / 2] = make_real2(x[n - 1], 0.0);
    }
    c.sync();
#pragma unroll
    for (int j = 0; j + 1 < n; j += 2) {
      const real2 v = slot[j / 2];
      out[j] = v.x;
      out[j + 1] = v.y;
    }
    if (n % 2) out[n - 1] = slot[n / 2].x;
  }
  // out[j] = x of lane j (j < n)
  template <int n>
  static __device__ __forceinline__ void allgather(Ctx& c, real x, real (&out)[n]) {
    real* v = c.sm + c.nmat * MAT + 4 * BCW + c.vflip * VEC;
    c.vflip ^= 1;
    if (c.r < n) v[c.r] = x;
    c.sync();
#pragma unroll
    for (int j = 0; j < n; ++j) out[j] = v[j];
  }
  // row `row` of matrix `which` <- x   (caller syncs)
  static __device__ __forceinline__ void put_row(Ctx& c, int which, int row, const real (&x)[D]) {
    real* M = c.mat(which) + row * LDM;
#pragma unroll
    for (int j = 0; j < D; ++j) M[j] = x[j];
  }
  // y[c] = sum_k a[k] * M[k][c]   (row vector times staged matrix)
  static __device__ __forceinline__ void row_times(const Ctx& c, int which, const real (&a)[D], real (&y)[D]) {
    const real* M = c.mat(which);
#pragma unroll
    for (int j = 0; j < D; ++j) y[j] = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
#pragma unroll
      for (int j = 0; j < D; ++j) y[j] = fma(a[k], M[k * LDM + j], y[j]);
    }
  }
  // y[c] = sum_k M1[k][col] * M2[k][c]   (column `col` of M1, transposed, times M2)
  static __device__ __forceinline__ void col_times(const Ctx& c, int which1, int col, int which2, real (&y)[D]) {
    const real* M1 = c.mat(which1);
    const real* M2 = c.mat(which2);
#pragma unroll
    for (int j = 0; j < D; ++j) y[j] = 0.0;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const real a = M1[k * LDM + col];
#pragma unroll
      for (int j = 0; j < D; ++j) y[j] = fma(a, M2[k * LDM + j], y[j]);
    }
  }
  // Operand loads go to L2 (ld.global.cg): inside a dataflow sweep (k_tree_flow) the operands were written by other
  // SMs during the SAME kernel, and a node's first/last cache line can be shared with its neighbour's, so a line
  // cached in L1 by an earlier read may be stale.  (Per-level launches read L2-resident data anyway.)
  static __device__ __forceinline__ real ldg(const real* p) { return __ldcg(p); }
  static __device__ __forceinline__ void load_row(const real* p, real (&x)[D]) {
    if constexpr (D % 2 == 0) {
      const real2* p2 = reinterpret_cast<const real2*>(p);
#pragma unroll
      for (int j = 0; j < D / 2; ++j) {
        const real2 v = __ldcg(p2 + j);
        x[2 * j] = v.x;
        x[2 * j + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < D; ++j) x[j] = __ldcg(p + j);
    }
  }
  static __device__ __forceinline__ void store_row_tri(real* __restrict__ p, int row, const real* x) {
#pragma unroll
    for (int j = 0; j < D; ++j) p[j] = (j <= row) ? x[j] : 0.0;
  }

  // one pivot of a right-Householder triangularisation over NC columns; rows over lanes; two independent halves
  // (rows [0,D) use pivot lane I, rows [D,2D) use pivot lane D+I) when DUAL, else a single matrix whose rows are the
  // lanes [0, NR)
  template <int I, int NC, int NPIV, bool DUAL>
  static __device__ __forceinline__ void tria_step(Ctx& cx, real (&x)[NC]) {
    if constexpr (I < NPIV && I + 1 < NC) {
      constexpr int n = NC - I;
      real mine[n], piv[n];
#pragma unroll
      for (int j = I; j < NC; ++j) mine[j - I] = x[j];
      const bool second = DUAL && cx.r >= D;
      if (DUAL) {
        // both halves publish into their own slot; each lane reads its half's slot
        real2* s0 = reinterpret_cast<real2*>(cx.sm + cx.nmat * MAT + (0 + cx.flip) * BCW);
        real2* s1 = reinterpret_cast<real2*>(cx.sm + cx.nmat * MAT + (2 + cx.flip) * BCW);
        if (cx.r == I || cx.r == D + I) {
          real2* s = second ? s1 : s0;
#pragma unroll
          for (int j = 0; j + 1 < n; j += 2) s[j / 2] = make_real2(mine[j], mine[j + 1]);
          if (n % 2) s[n / 2] = make_real2(mine[n - 1], 0.0);
        }
        cx.sync();
        const real2* s = second ? s1 : s0;
#pragma unroll
        for (int j = 0; j + 1 < n; j += 2) {
          const real2 v = s[j / 2];
          piv[j] = v.x;
          piv[j + 1] = v.y;
        }
        if (n % 2) piv[n - 1] = s[n / 2].x;
        cx.flip ^= 1;
      } else {
        bcast<n>(cx, 0, mine, I, piv);
        cx.flip ^= 1;
      }
      const HH h = house<n - 1>(piv[0], piv + 1);
      real w = fma(h.s, x[I], dotn<NC - I - 1>(&x[I + 1], piv + 1));  // the dot does not wait for the reflector
      const int rr = second ? cx.r - D : cx.r;
      w = (rr >= I) ? w * h.tp : 0.0;
      x[I] = (rr == I) ? h.beta : fma(-w, h.s, x[I]);
#pragma unroll
      for (int j = I + 1; j < NC; ++j) x[j] = fma(-w, piv[j - I], x[j]);
      tria_step<I + 1, NC, NPIV, DUAL>(cx, x);
    }
  }

  // ------------------------------------------------------------------------------------------- filtering operator
  // e1: earlier (packed element, or packed state when STATE), e2: later element; out: element, or state when STATE
  template <bool STATE>
  static __device__ __forceinline__ void filter_combine(Ctx& cx, const real* __restrict__ e1,
                                                        const real* __restrict__ e2, real* __restrict__ out) {
    constexpr int DD = D * D;
    const int r = cx.r;
    const bool top = r < D;
    const bool bot = r >= D && r < W2;
    const int rr = top ? r : (bot ? r - D : 0);
    // element pointers
    const real* A1 = e1;
    const real* b1 = STATE ? e1 : e1 + DD;
    const real* U1 = STATE ? e1 + D : e1 + DD + D;
    const real* n1 = e1 + 2 * DD + D;
    const real* Z1 = e1 + 2 * DD + 2 * D;
    const real* A2 = e2;
    const real* b2 = e2 + DD;
    const real* U2 = e2 + DD + D;
    const real* n2 = e2 + 2 * DD + D;
    const real* Z2 = e2 + 2 * DD + 2 * D;

    // ALL global operands are loaded here, in one round of independent loads: every later load would sit behind a
    // __syncwarp (a memory barrier the compiler cannot hoist loads across) and cost its own L2 round trip (~0.4 us)
    // on the dependent chain of the sweep -- five such round trips per combine before this was hoisted.
    real u1[D], z2[D], a1[D], a2[D], q2[D];
    load_row(U1 + rr * D, u1);
    load_row(Z2 + rr * D, z2);
    if (!STATE) load_row(A1 + rr * D, a1);
    load_row(A2 + rr * D, a2);
    load_row((bot && !STATE) ? Z1 + rr * D : U2 + rr * D, q2);
    const real b1r = ldg(b1 + rr), n2r = ldg(n2 + rr), b2r = ldg(b2 + rr);
    real n1r = 0.0;
    if (!STATE) n1r = ldg(n1 + rr);
    // stage U1, Z2 (and A1): lanes [0,D) publish U1 rows, lanes [D,2D) publish Z2 rows
    if (top) put_row(cx, mU1, rr, u1);
    if (bot) put_row(cx, mZ2, rr, z2);
    if (!STATE) {
      if (top) put_row(cx, mA1, rr, a1);
    }
    cx.sync();
    // Xi rows: top r: [ (U1^T Z2)[r,:], e_r ] ; bottom r: [ Z2[r,:], 0 ]
    real x[W2];
    {
      real y[D];
      col_times(cx, mU1, rr, mZ2, y);
#pragma unroll
      for (int j = 0; j < D; ++j) {
        x[j] = top ? y[j] : (bot ? z2[j] : 0.0);
        x[D + j] = (top && j == rr) ? 1.0 : 0.0;
      }
    }
    // Only the first D pivots: they give Xi11 and Xi21.  The remaining block B (bottom rows, columns D..2D) is NOT
    // triangularised: it enters the result only through Z = tria([A1^T Xi22 | Z1]), which depends on Xi22 Xi22^T = B B^T
    // alone, so B itself serves (D-1 fewer pivots on the latency chain of every up-sweep level).
    tria_step<0, W2, D, false>(cx, x);
    // publish Xi11 (lanes top), Xi21 and B (lanes bottom)
    {
      real lo[D], hi[D];
#pragma unroll
      for (int j = 0; j < D; ++j) {
        lo[j] = x[j];
        hi[j] = x[D + j];
      }
      if (top) {
#pragma unroll
        for (int j = 0; j < D; ++j) lo[j] = (j <= rr) ? lo[j] : 0.0;
        put_row(cx, mX11, rr, lo);
      }
      if (bot) {
        put_row(cx, mX21, rr, lo);
        if (!STATE) put_row(cx, mX22, rr, hi);
      }
    }
    cx.sync();
    // Y row r = U1[r,:] Xi11^{-T}  (forward substitution), lanes top (others compute harmlessly on row rr)
    real y[D];
    {
      const real* X11 = cx.mat(mX11);
#pragma unroll
      for (int j = 0; j < D; ++j) {
        real acc = u1[j];
#pragma unroll
        for (int i = 0; i < j; ++i) acc = fma(-y[i], X11[j * LDM + i], acc);
        y[j] = acc * fast_rcp(X11[j * LDM + j]);
      }
    }
    if (top) put_row(cx, mY, rr, y);
    // G row r = e_r - Y[r,:] Xi21^T
    real g[D];
    {
      const real* X21 = cx.mat(mX21);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        real acc = (c == rr) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) acc = fma(-y[k], X21[c * LDM + k], acc);
        g[c] = acc;
      }
    }
    if (top) put_row(cx, mG, rr, g);
    cx.sync();
    // ---- b = A2 G (b1 + U1 U1^T eta2) + b2
    {
      real v[D], t[D];
      allgather<D>(cx, top ? n2r : 0.0, v);
      // (U1^T eta2)[rr]: column rr of U1
      real s = 0.0;
      const real* MU = cx.mat(mU1);
#pragma unroll
      for (int k = 0; k < D; ++k) s = fma(MU[k * LDM + rr], v[k], s);
      allgather<D>(cx, s, t);
      real t0 = b1r;
#pragma unroll
      for (int k = 0; k < D; ++k) t0 = fma(u1[k], t[k], t0);
      allgather<D>(cx, t0, v);
      real t2 = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) t2 = fma(g[k], v[k], t2);
      allgather<D>(cx, t2, t);
      real bo = b2r;
#pragma unroll
      for (int k = 0; k < D; ++k) bo = fma(a2[k], t[k], bo);
      if (top) (STATE ? out : out + DD)[rr] = bo;
    }
    // ---- rows for the second triangularisation: top: [A2 Y | U2] -> U ; bottom: [A1^T Xi22 | Z1] -> Z
    real x2[W2];
    {
      real p[D];
      row_times(cx, mY, a2, p);
      if (!STATE) {
        real pz[D];
        col_times(cx, mA1, rr, mX22, pz);
#pragma unroll
        for (int j = 0; j < D; ++j) p[j] = bot ? pz[j] : p[j];
      }
#pragma unroll
      for (int j = 0; j < D; ++j) {
        x2[j] = p[j];
        x2[D + j] = q2[j];
      }
    }
    if (!STATE) {
      // ---- A = A2 (G A1)
      real pr[D], ao[D];
      row_times(cx, mA1, g, pr);
      if (top) put_row(cx, mP, rr, pr);
      cx.sync();
      row_times(cx, mP, a2, ao);
      if (top) {
#pragma unroll
        for (int j = 0; j < D; ++j) out[rr * D + j] = ao[j];
      }
      // ---- eta = A1^T G^T (eta2 - Z2 Z2^T b1) + eta1
      real v[D], t[D];
      allgather<D>(cx, top ? b1r : 0.0, v);
      real s = 0.0;
      const real* MZ = cx.mat(mZ2);
#pragma unroll
      for (int k = 0; k < D; ++k) s = fma(MZ[k * LDM + rr], v[k], s);  // (Z2^T b1)[rr]
      allgather<D>(cx, s, t);
      real s0 = n2r;
#pragma unroll
      for (int k = 0; k < D; ++k) s0 = fma(-MZ[rr * LDM + k], t[k], s0);  // eta2 - Z2 (Z2^T b1)
      allgather<D>(cx, s0, v);
      real s2 = 0.0;
      const real* MG = cx.mat(mG);
#pragma unroll
      for (int k = 0; k < D; ++k) s2 = fma(MG[k * LDM + rr], v[k], s2);  // (G^T s)[rr]
      allgather<D>(cx, s2, t);
      real eo = n1r;
      const real* MA = cx.mat(mA1);
#pragma unroll
      for (int k = 0; k < D; ++k) eo = fma(MA[k * LDM + rr], t[k], eo);  // (A1^T .)[rr]
      if (top) out[2 * DD + D + rr] = eo;
    }
    tria_step<0, W2, D, !STATE>(cx, x2);
    if (top) store_row_tri((STATE ? out + D : out + DD + D) + rr * D, rr, x2);
    if (!STATE && bot) store_row_tri(out + 2 * DD + 2 * D + rr * D, rr, x2);
  }

  // ------------------------------------------------------------------------------- chunk-level smoothing element
  // The smoothing element (g, E, Dm) of a whole chunk (steps s..e), i.e. the composition of its per-step backward
  // kernels (smoother.py:37-63), obtained directly from the filtered state (m, L) at the chunk start and the chunk's
  // filtering element taken BEFORE its last measurement update (A, b, U, eta, Z):
  //    x_s | y_{1:e-1} ~ N(m', Y Y^T),  Y = L Xi11^{-T},  m' = G (m + L L^T eta)      (as in the filtering operator)
  //    tria([[A Y, U],[Y, 0]]) = [[Phi11, 0],[Phi21, Phi22]],  E = Phi21 Phi11^{-1},  g = m' - E (A m' + b),  Dm = Phi22
  // One such op per chunk replaces L-1 per-step compositions inside the filter scan.
  static __device__ __forceinline__ void chunk_kernel(Ctx& cx, const real* __restrict__ st,
                                                      const real* __restrict__ e2, real* __restrict__ out) {
    constexpr int DD = D * D;
    const int r = cx.r;
    const bool top = r < D;
    const bool bot = r >= D && r < W2;
    const int rr = top ? r : (bot ? r - D : 0);
    const real* b1 = st;
    const real* U1 = st + D;
    const real* A2 = e2;
    const real* b2 = e2 + DD;
    const real* U2 = e2 + DD + D;
    const real* n2 = e2 + 2 * DD + D;
    const real* Z2 = e2 + 2 * DD + 2 * D;
    real u1[D], z2[D], a2[D], q2[D];
    load_row(U1 + rr * D, u1);
    load_row(Z2 + rr * D, z2);
    load_row(A2 + rr * D, a2);
    load_row(U2 + rr * D, q2);
    const real b1r = ldg(b1 + rr), n2r = ldg(n2 + rr), b2r = ldg(b2 + rr);
    if (top) put_row(cx, mU1, rr, u1);
    if (bot) put_row(cx, mZ2, rr, z2);
    cx.sync();
    real x[W2];
    {
      real y0[D];
      col_times(cx, mU1, rr, mZ2, y0);
#pragma unroll
      for (int j = 0; j < D; ++j) {
        x[j] = top ? y0[j] : (bot ? z2[j] : 0.0);
        x[D + j] = (top && j == rr) ? 1.0 : 0.0;
      }
    }
    tria_step<0, W2, D, false>(cx, x);
    {
      real lo[D];
#pragma unroll
      for (int j = 0; j < D; ++j) lo[j] = (top && j > rr) ? 0.0 : x[j];
      if (top) put_row(cx, mX11, rr, lo);
      if (bot) put_row(cx, mX21, rr, lo);
    }
    cx.sync();
    real y[D];
    {
      const real* X11 = cx.mat(mX11);
#pragma unroll
      for (int j = 0; j < D; ++j) {
        real acc = u1[j];
#pragma unroll
        for (int i = 0; i < j; ++i) acc = fma(-y[i], X11[j * LDM + i], acc);
        y[j] = acc * fast_rcp(X11[j * LDM + j]);
      }
    }
    if (top) put_row(cx, mY, rr, y);
    real g[D];
    {
      const real* X21 = cx.mat(mX21);
#pragma unroll
      for (int c = 0; c < D; ++c) {
        real acc = (c == rr) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 0; k < D; ++k) acc = fma(-y[k], X21[c * LDM + k], acc);
        g[c] = acc;
      }
    }
    cx.sync();
    // m' = G (m + L L^T eta)  and  v = A m' + b
    real mp[D], vf[D];
    {
      real v[D], t[D];
      allgather<D>(cx, top ? n2r : 0.0, v);
      real s = 0.0;
      const real* MU = cx.mat(mU1);
#pragma unroll
      for (int k = 0; k < D; ++k) s = fma(MU[k * LDM + rr], v[k], s);
      allgather<D>(cx, s, t);
      real t0 = b1r;
#pragma unroll
      for (int k = 0; k < D; ++k) t0 = fma(u1[k], t[k], t0);
      allgather<D>(cx, t0, v);
      real t2 = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) t2 = fma(g[k], v[k], t2);
      allgather<D>(cx, t2, mp);
      real av = b2r;
#pragma unroll
      for (int k = 0; k < D; ++k) av = fma(a2[k], mp[k], av);
      allgather<D>(cx, av, vf);
    }
    // joint array [[A Y, U],[Y, 0]]
    real x2[W2];
    {
      real p[D];
      row_times(cx, mY, a2, p);
#pragma unroll
      for (int j = 0; j < D; ++j) {
        x2[j] = top ? p[j] : (bot ? y[j] : 0.0);
        x2[D + j] = top ? q2[j] : 0.0;
      }
    }
    tria_step<0, W2, W2, false>(cx, x2);
    {
      real lo[D];
#pragma unroll
      for (int j = 0; j < D; ++j) lo[j] = (j <= rr) ? x2[j] : 0.0;
      cx.sync();
      if (top) put_row(cx, mX11, rr, lo);
      cx.sync();
    }
    // E row (bottom lanes): e Phi11 = phi21  (backward substitution over the columns)
    real e[D];
    {
      const real* P11 = cx.mat(mX11);
#pragma unroll
      for (int j = D - 1; j >= 0; --j) {
        real acc = x2[j];
#pragma unroll
        for (int i = j + 1; i < D; ++i) acc = fma(-e[i], P11[i * LDM + j], acc);
        e[j] = acc * fast_rcp(P11[j * LDM + j]);
      }
    }
    if (bot) {
      real go = 0.0;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        go = (k == rr) ? mp[k] : go;
      }
#pragma unroll
      for (int k = 0; k < D; ++k) go = fma(-e[k], vf[k], go);
      out[rr] = go;
#pragma unroll
      for (int j = 0; j < D; ++j) out[D + rr * D + j] = e[j];
      store_row_tri(out + D + DD + rr * D, rr, x2 + D);
    }
  }

  // ------------------------------------------------------------------------------------------- smoothing operator
  // e1: LATER (packed element, or packed state when STATE), e2: EARLIER element.  GS lanes, lane r owns row r.
  template <bool STATE>
  static __device__ __forceinline__ void smooth_combine(Ctx& cx, const real* __restrict__ e1,
                                                        const real* __restrict__ e2, real* __restrict__ out) {
    constexpr int DD = D * D;
    const int r = cx.r;
    const bool act = r < D;
    const int rr = act ? r : 0;
    const real* g1 = e1;
    const real* E1 = e1 + D;
    const real* D1 = STATE ? e1 + D : e1 + D + DD;
    const real* g2 = e2;
    const real* E2 = e2 + D;
    const real* D2 = e2 + D + DD;
    real row[D], row1[D], e2r[D], d2[D], v[D];
    load_row(D1 + rr * D, row);
    if (!STATE) load_row(E1 + rr * D, row1);
    load_row(E2 + rr * D, e2r);
    load_row(D2 + rr * D, d2);
    const real g1r = ldg(g1 + rr), g2r = ldg(g2 + rr);
    if (act) put_row(cx, sD1, rr, row);
    if (!STATE) {
      if (act) put_row(cx, sE1, rr, row1);
    }
    allgather<D>(cx, act ? g1r : 0.0, v);  // also orders the put_rows before the reads below
    real go = g2r;
#pragma unroll
    for (int k = 0; k < D; ++k) go = fma(e2r[k], v[k], go);
    if (act) out[rr] = go;
    if (!STATE) {
      real eo[D];
      row_times(cx, sE1, e2r, eo);
      if (act) {
#pragma unroll
        for (int j = 0; j < D; ++j) out[D + rr * D + j] = eo[j];
      }
    }
    real x[W2];
    {
      real p[D];
      row_times(cx, sD1, e2r, p);
#pragma unroll
      for (int j = 0; j < D; ++j) {
        x[j] = p[j];
        x[D + j] = d2[j];
      }
    }
    tria_step<0, W2, D, false>(cx, x);
    if (act) store_row_tri((STATE ? out + D : out + D + DD) + rr * D, rr, x);
  }
};

}  // namespace POF_NS
