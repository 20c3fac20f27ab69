// Carry level of the blocked scans: the reference's associative operators on *general* elements, one warp per
// element pair, matrices staged in shared memory, Householder triangularisation with the rows spread over the
// lanes.  Runtime state dimension D (any d, q).
//
//   filter_combine   sqrt_filtering_operator  (pof/parallel_filtsmooth/filter.py:117-142)
//   filter_apply     the same operator with elem1 = a *state* (A=0, b=m, U=L, eta=0, Z=0), producing only (b, U):
//                    what the exclusive down-sweep needs (filtered state at a chunk boundary)
//   smooth_combine   sqrt_smoothing_operator  (pof/parallel_filtsmooth/smoother.py:53-63)
//   smooth_apply     the same with elem1 = a state (g=m, E=0, D=L), producing (m, L)
//
// Algebra used (equivalent to the reference formulas; Xi11, Xi21, Xi22 as in filter.py:125-129):
//   Y  = U1 Xi11^{-T}            V = Y Xi21^T            G = I - V
//   A  = A2 G A1                 b = A2 G (b1 + U1 U1^T eta2) + b2        U = tria([A2 Y, U2])
//   eta = A1^T G^T (eta2 - Z2 Z2^T b1) + eta1                              Z = tria([A1^T Xi22, Z1])
// because  M^T = (Xi11^{-1} U1^T A2^T)^T = A2 Y,  m^T Xi21^T = V,  _e^T U1^T = V^T.
//
// Packed element layouts (doubles):  filter  [A D*D | b D | U D*D | eta D | Z D*D]   (3 D^2 + 2 D)
//                                    smoother [g D | E D*D | Dm D*D]                  (2 D^2 + D)
//                                    state    [m D | L D*D]                           (D^2 + D)
// The code also compiles for the host with a 1-lane "warp" (tests/hostsim).
#pragma once
#include "pof_small.cuh"

namespace pof {

#if defined(__CUDACC__)
#define POF_DEV __host__ __device__ __forceinline__
#else
#define POF_DEV inline
#endif

// A "warp" of cooperating lanes: 32 on the device, a single lane in the host simulator.
struct Warp {
  int lane;
#if defined(__CUDA_ARCH__)
  static constexpr int NL = 32;
  __device__ __forceinline__ Warp() : lane(threadIdx.x & 31) {}
  __device__ __forceinline__ void sync() const { __syncwarp(); }
  __device__ __forceinline__ double sum(double x) const {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
  }
#else
  static constexpr int NL = 1;
  Warp() : lane(0) {}
  void sync() const {}
  double sum(double x) const { return x; }
#endif
};

POF_DEV int filter_elem_size(int D) { return 3 * D * D + 2 * D; }
POF_DEV int smooth_elem_size(int D) { return 2 * D * D + D; }
POF_DEV int state_size(int D) { return D * D + D; }

// in-place right-Householder lower-triangularisation of M (R x C, leading dimension ld), pivots 0..npiv-1.
// The tails of the pivot rows hold the (dead) reflector vectors afterwards; only the lower triangle is meaningful.
POF_DEV void coop_tria(const Warp& w, double* M, int R, int C, int ld, int npiv) {
  for (int i = 0; i < npiv && i + 1 < C; ++i) {
    double* pr = M + i * ld;
    double s = 0.0;
    for (int j = i + 1 + w.lane; j < C; j += Warp::NL) s = fma(pr[j], pr[j], s);
    s = w.sum(s);
    const double alpha = pr[i];
    const bool nz = s > 0.0;
    const double nrm = sqrt(fma(alpha, alpha, s));
    const double beta = (alpha >= 0.0) ? -nrm : nrm;
    const double tau = nz ? (beta - alpha) / (nz ? beta : 1.0) : 0.0;
    const double scale = nz ? 1.0 / (alpha - beta) : 0.0;
    w.sync();
    for (int j = i + 1 + w.lane; j < C; j += Warp::NL) pr[j] *= scale;
    if (w.lane == 0) pr[i] = nz ? beta : alpha;
    w.sync();
    if (nz) {
      for (int r = i + 1 + w.lane; r < R; r += Warp::NL) {
        double* row = M + r * ld;
        double t = row[i];
        for (int j = i + 1; j < C; ++j) t = fma(row[j], pr[j], t);
        t *= tau;
        row[i] -= t;
        for (int j = i + 1; j < C; ++j) row[j] = fma(-t, pr[j], row[j]);
      }
    }
    w.sync();
  }
}

// out(r,c) = sum_k X(r,k) * Y(k,c), r<R, c<Cn, k<K, via accessor lambdas; lanes over output entries
template <class FX, class FY, class FO>
POF_DEV void coop_gemm(const Warp& w, int R, int Cn, int K, FX fx, FY fy, FO fo) {
  for (int idx = w.lane; idx < R * Cn; idx += Warp::NL) {
    const int r = idx / Cn, c = idx - r * Cn;
    double s = 0.0;
    for (int k = 0; k < K; ++k) s = fma(fx(r, k), fy(k, c), s);
    fo(r, c, s);
  }
}

// Shared-memory footprint (doubles) of one warp's workspace
POF_DEV int coop_ws_doubles(int D) {
  const int ldx = 2 * D + 1;
  return 6 * D * D + 4 * D + 2 * D * ldx + 3 * D * D + D * ldx + 4 * D;
}

struct CoopWs {
  double *A1, *U1, *Z1, *A2, *U2, *Z2, *b1, *e1, *b2, *e2;
  double* Xi;   // 2D x 2D, ld = 2D+1
  double *Y, *G, *P;  // D x D
  double* W;    // D x 2D, ld = 2D+1
  double *t0, *t1, *t2, *t3;
  int ldx;
  POF_DEV CoopWs(double* base, int D) {
    ldx = 2 * D + 1;
    const int DD = D * D;
    A1 = base; U1 = A1 + DD; Z1 = U1 + DD; A2 = Z1 + DD; U2 = A2 + DD; Z2 = U2 + DD;
    b1 = Z2 + DD; e1 = b1 + D; b2 = e1 + D; e2 = b2 + D;
    Xi = e2 + D;
    Y = Xi + 2 * D * ldx; G = Y + DD; P = G + DD;
    W = P + DD;
    t0 = W + D * ldx; t1 = t0 + D; t2 = t1 + D; t3 = t2 + D;
  }
};

POF_DEV void coop_copy(const Warp& w, double* dst, const double* src, int n) {
  for (int i = w.lane; i < n; i += Warp::NL) dst[i] = src[i];
}
POF_DEV void coop_zero(const Warp& w, double* dst, int n) {
  for (int i = w.lane; i < n; i += Warp::NL) dst[i] = 0.0;
}

// The filtering operator.  e1: earlier element (packed filter element, or packed state if state_mode),
// e2: later element (packed filter element).  out: packed filter element, or packed state if state_mode.
POF_DEV void filter_combine(const Warp& w, int D, const double* e1, const double* e2, double* out, double* smem,
                            bool state_mode) {
  CoopWs s(smem, D);
  const int DD = D * D, ldx = s.ldx;
  // ---- stage operands
  if (state_mode) {
    coop_copy(w, s.b1, e1, D);
    coop_copy(w, s.U1, e1 + D, DD);
  } else {
    coop_copy(w, s.A1, e1, DD);
    coop_copy(w, s.b1, e1 + DD, D);
    coop_copy(w, s.U1, e1 + DD + D, DD);
    coop_copy(w, s.e1, e1 + 2 * DD + D, D);
    coop_copy(w, s.Z1, e1 + 2 * DD + 2 * D, DD);
  }
  coop_copy(w, s.A2, e2, DD);
  coop_copy(w, s.b2, e2 + DD, D);
  coop_copy(w, s.U2, e2 + DD + D, DD);
  coop_copy(w, s.e2, e2 + 2 * DD + D, D);
  coop_copy(w, s.Z2, e2 + 2 * DD + 2 * D, DD);
  w.sync();
  // ---- Xi = [[U1^T Z2, I],[Z2, 0]]
  double* Xi = s.Xi;
  coop_gemm(w, D, D, D, [&](int r, int k) { return s.U1[k * D + r]; }, [&](int k, int c) { return s.Z2[k * D + c]; },
            [&](int r, int c, double v) { Xi[r * ldx + c] = v; });
  for (int idx = w.lane; idx < DD; idx += Warp::NL) {
    const int r = idx / D, c = idx - r * D;
    Xi[r * ldx + D + c] = (r == c) ? 1.0 : 0.0;
    Xi[(D + r) * ldx + c] = s.Z2[idx];
    Xi[(D + r) * ldx + D + c] = 0.0;
  }
  w.sync();
  coop_tria(w, Xi, 2 * D, 2 * D, ldx, state_mode ? D : 2 * D);
  // ---- Y = U1 Xi11^{-T}: row r of Y solves  y Xi11^T = u_r  (forward in j); lanes over rows
  for (int r = w.lane; r < D; r += Warp::NL) {
    for (int j = 0; j < D; ++j) {
      double acc = s.U1[r * D + j];
      for (int i = 0; i < j; ++i) acc = fma(-s.Y[r * D + i], Xi[j * ldx + i], acc);
      s.Y[r * D + j] = acc / Xi[j * ldx + j];
    }
  }
  w.sync();
  // ---- G = I - Y Xi21^T
  coop_gemm(w, D, D, D, [&](int r, int k) { return s.Y[r * D + k]; },
            [&](int k, int c) { return Xi[(D + c) * ldx + k]; },
            [&](int r, int c, double v) { s.G[r * D + c] = ((r == c) ? 1.0 : 0.0) - v; });
  // ---- t0 = b1 + U1 (U1^T eta2)
  for (int i = w.lane; i < D; i += Warp::NL) {
    double acc = 0.0;
    for (int k = 0; k < D; ++k) acc = fma(s.U1[k * D + i], s.e2[k], acc);
    s.t1[i] = acc;
  }
  w.sync();
  for (int i = w.lane; i < D; i += Warp::NL) {
    double acc = s.b1[i];
    for (int k = 0; k < D; ++k) acc = fma(s.U1[i * D + k], s.t1[k], acc);
    s.t0[i] = acc;
  }
  w.sync();
  // ---- t2 = G t0 ; b = A2 t2 + b2
  for (int i = w.lane; i < D; i += Warp::NL) {
    double acc = 0.0;
    for (int k = 0; k < D; ++k) acc = fma(s.G[i * D + k], s.t0[k], acc);
    s.t2[i] = acc;
  }
  w.sync();
  double* ob = state_mode ? out : out + DD;
  for (int i = w.lane; i < D; i += Warp::NL) {
    double acc = s.b2[i];
    for (int k = 0; k < D; ++k) acc = fma(s.A2[i * D + k], s.t2[k], acc);
    ob[i] = acc;
  }
  // ---- U = tria([A2 Y, U2])
  coop_gemm(w, D, D, D, [&](int r, int k) { return s.A2[r * D + k]; }, [&](int k, int c) { return s.Y[k * D + c]; },
            [&](int r, int c, double v) { s.W[r * ldx + c] = v; });
  for (int idx = w.lane; idx < DD; idx += Warp::NL) {
    const int r = idx / D, c = idx - r * D;
    s.W[r * ldx + D + c] = s.U2[idx];
  }
  w.sync();
  coop_tria(w, s.W, D, 2 * D, ldx, D);
  double* oU = state_mode ? out + D : out + DD + D;
  for (int idx = w.lane; idx < DD; idx += Warp::NL) {
    const int r = idx / D, c = idx - r * D;
    oU[idx] = (c <= r) ? s.W[r * ldx + c] : 0.0;
  }
  if (state_mode) return;
  w.sync();
  // ---- A = A2 (G A1)
  coop_gemm(w, D, D, D, [&](int r, int k) { return s.G[r * D + k]; }, [&](int k, int c) { return s.A1[k * D + c]; },
            [&](int r, int c, double v) { s.P[r * D + c] = v; });
  w.sync();
  coop_gemm(w, D, D, D, [&](int r, int k) { return s.A2[r * D + k]; }, [&](int k, int c) { return s.P[k * D + c]; },
            [&](int r, int c, double v) { out[r * D + c] = v; });
  // ---- eta = A1^T G^T (eta2 - Z2 Z2^T b1) + eta1
  for (int i = w.lane; i < D; i += Warp::NL) {
    double acc = 0.0;
    for (int k = 0; k < D; ++k) acc = fma(s.Z2[k * D + i], s.b1[k], acc);
    s.t1[i] = acc;
  }
  w.sync();
  for (int i = w.lane; i < D; i += Warp::NL) {
    double acc = s.e2[i];
    for (int k = 0; k < D; ++k) acc = fma(-s.Z2[i * D + k], s.t1[k], acc);
    s.t0[i] = acc;
  }
  w.sync();
  for (int i = w.lane; i < D; i += Warp::NL) {
    double acc = 0.0;
    for (int k = 0; k < D; ++k) acc = fma(s.G[k * D + i], s.t0[k], acc);
    s.t2[i] = acc;
  }
  w.sync();
  for (int i = w.lane; i < D; i += Warp::NL) {
    double acc = s.e1[i];
    for (int k = 0; k < D; ++k) acc = fma(s.A1[k * D + i], s.t2[k], acc);
    out[2 * DD + D + i] = acc;
  }
  // ---- Z = tria([A1^T Xi22, Z1])   (Xi22 lower triangular)
  coop_gemm(w, D, D, D, [&](int r, int k) { return s.A1[k * D + r]; },
            [&](int k, int c) { return (c <= k) ? Xi[(D + k) * ldx + D + c] : 0.0; },
            [&](int r, int c, double v) { s.W[r * ldx + c] = v; });
  for (int idx = w.lane; idx < DD; idx += Warp::NL) {
    const int r = idx / D, c = idx - r * D;
    s.W[r * ldx + D + c] = s.Z1[idx];
  }
  w.sync();
  coop_tria(w, s.W, D, 2 * D, ldx, D);
  double* oZ = out + 2 * DD + 2 * D;
  for (int idx = w.lane; idx < DD; idx += Warp::NL) {
    const int r = idx / D, c = idx - r * D;
    oZ[idx] = (c <= r) ? s.W[r * ldx + c] : 0.0;
  }
}

// The smoothing operator.  e1: LATER element (packed smoother element, or packed state if state_mode),
// e2: EARLIER element.  g = E2 g1 + g2 ; E = E2 E1 ; Dm = tria([E2 D1, D2]).
POF_DEV void smooth_combine(const Warp& w, int D, const double* e1, const double* e2, double* out, double* smem,
                            bool state_mode) {
  CoopWs s(smem, D);
  const int DD = D * D, ldx = s.ldx;
  const double* g1 = e1;
  const double* D1 = state_mode ? e1 + D : e1 + D + DD;
  coop_copy(w, s.b1, g1, D);
  coop_copy(w, s.U1, D1, DD);
  if (!state_mode) coop_copy(w, s.A1, e1 + D, DD);
  coop_copy(w, s.b2, e2, D);
  coop_copy(w, s.A2, e2 + D, DD);
  coop_copy(w, s.U2, e2 + D + DD, DD);
  w.sync();
  for (int i = w.lane; i < D; i += Warp::NL) {
    double acc = s.b2[i];
    for (int k = 0; k < D; ++k) acc = fma(s.A2[i * D + k], s.b1[k], acc);
    out[i] = acc;
  }
  if (!state_mode) {
    coop_gemm(w, D, D, D, [&](int r, int k) { return s.A2[r * D + k]; },
              [&](int k, int c) { return s.A1[k * D + c]; }, [&](int r, int c, double v) { out[D + r * D + c] = v; });
  }
  coop_gemm(w, D, D, D, [&](int r, int k) { return s.A2[r * D + k]; }, [&](int k, int c) { return s.U1[k * D + c]; },
            [&](int r, int c, double v) { s.W[r * ldx + c] = v; });
  for (int idx = w.lane; idx < DD; idx += Warp::NL) {
    const int r = idx / D, c = idx - r * D;
    s.W[r * ldx + D + c] = s.U2[idx];
  }
  w.sync();
  coop_tria(w, s.W, D, 2 * D, ldx, D);
  double* oD = state_mode ? out + D : out + D + DD;
  for (int idx = w.lane; idx < DD; idx += Warp::NL) {
    const int r = idx / D, c = idx - r * D;
    oD[idx] = (c <= r) ? s.W[r * ldx + c] : 0.0;
  }
}

}  // namespace pof
