// Thread-private small dense linear algebra for the leaf level of the blocked IEKS scans.
//
// Everything here works on fixed-size C arrays with compile-time loop bounds so that, for small state
// dimension D, nvcc fully unrolls the loops and keeps the matrices in registers (no shuffles, no shared
// memory: one thread owns one time-chunk).  The same code compiles for the host (tests/hostsim) so the
// math is validated against the CPU oracle without a GPU.
//
// The orthogonal triangularisations replace the reference's `tria(A) = qr(A^T, mode="r")^T`
// (pof/utils.py:33-41): Householder reflections applied from the right, LAPACK dlarfg sign convention
// (beta = -sign(alpha) * norm), never forming Q.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define POF_HD __host__ __device__ __forceinline__
#else
#define POF_HD inline
#endif

// `#pragma unroll` with a compile-time factor: full unroll for small D, rolled loops otherwise.
#define POF_PRAGMA(x) _Pragma(#x)
#if defined(__CUDACC__)
#define POF_UNROLL_N(n) POF_PRAGMA(unroll n)
#else
#define POF_UNROLL_N(n)
#endif

namespace pof {

// Householder generator for the row vector (alpha, x[0..n-1]); returns tau, overwrites alpha with beta and
// x with v (v0 = 1 implicit).  Zero tail => tau = 0 (identity), no NaN for all-zero rows.
template <int N, int UF>
POF_HD double house_gen(double& alpha, double* x) {
  double sigma = 0.0;
  POF_UNROLL_N(UF)
  for (int j = 0; j < N; ++j) sigma = fma(x[j], x[j], sigma);
  const bool nz = sigma > 0.0;
  const double nrm = sqrt(fma(alpha, alpha, sigma));
  const double beta = (alpha >= 0.0) ? -nrm : nrm;
  const double den = nz ? beta : 1.0;
  const double tau = nz ? (beta - alpha) / den : 0.0;
  const double scale = nz ? 1.0 / (alpha - beta) : 0.0;
  POF_UNROLL_N(UF)
  for (int j = 0; j < N; ++j) x[j] *= scale;
  alpha = nz ? beta : alpha;
  return tau;
}

// Triangular-pentagonal right-QR.
//   pivot rows      [ T | C ]   T: NP x NP lower triangular (entries above the diagonal never referenced),
//                               C: NP x K dense
//   passenger rows  [ PT | PC ] PT: NB x NP (column i is touched by pivot i only), PC: NB x K
// After the call  [T|C] Q = [T'|0]  and  [PT|PC] Q = [PT'|PC'] for the same orthogonal Q.
template <int NP, int K, int NB, int UF>
POF_HD void tpqrt(double (&T)[NP][NP], double (&C)[NP][K], double (*PT)[NP], double (*PC)[K]) {
  POF_UNROLL_N(UF)
  for (int i = 0; i < NP; ++i) {
    const double tau = house_gen<K, UF>(T[i][i], C[i]);
    POF_UNROLL_N(UF)
    for (int r = i + 1; r < NP; ++r) {
      double w = T[r][i];
      POF_UNROLL_N(UF)
      for (int j = 0; j < K; ++j) w = fma(C[r][j], C[i][j], w);
      w *= tau;
      T[r][i] -= w;
      POF_UNROLL_N(UF)
      for (int j = 0; j < K; ++j) C[r][j] = fma(-w, C[i][j], C[r][j]);
    }
    POF_UNROLL_N(UF)
    for (int r = 0; r < NB; ++r) {
      double w = PT[r][i];
      POF_UNROLL_N(UF)
      for (int j = 0; j < K; ++j) w = fma(PC[r][j], C[i][j], w);
      w *= tau;
      PT[r][i] -= w;
      POF_UNROLL_N(UF)
      for (int j = 0; j < K; ++j) PC[r][j] = fma(-w, C[i][j], PC[r][j]);
    }
  }
}

// Plain right-Householder lower-triangularisation of X (R x Cc), pivots 0..NPIV-1.  Entries right of the
// diagonal in the pivot rows are set to zero.
template <int R, int Cc, int NPIV, int UF>
POF_HD void house_rows(double (&X)[R][Cc]) {
  POF_UNROLL_N(UF)
  for (int i = 0; i < NPIV; ++i) {
    if (i + 1 >= Cc) break;
    double v[Cc];
    double sigma = 0.0;
    POF_UNROLL_N(UF)
    for (int j = i + 1; j < Cc; ++j) sigma = fma(X[i][j], X[i][j], sigma);
    const bool nz = sigma > 0.0;
    const double alpha = X[i][i];
    const double nrm = sqrt(fma(alpha, alpha, sigma));
    const double beta = (alpha >= 0.0) ? -nrm : nrm;
    const double tau = nz ? (beta - alpha) / (nz ? beta : 1.0) : 0.0;
    const double scale = nz ? 1.0 / (alpha - beta) : 0.0;
    POF_UNROLL_N(UF)
    for (int j = i + 1; j < Cc; ++j) {
      v[j] = X[i][j] * scale;
      X[i][j] = 0.0;
    }
    X[i][i] = nz ? beta : alpha;
    POF_UNROLL_N(UF)
    for (int r = i + 1; r < R; ++r) {
      double w = X[r][i];
      POF_UNROLL_N(UF)
      for (int j = i + 1; j < Cc; ++j) w = fma(X[r][j], v[j], w);
      w *= tau;
      X[r][i] -= w;
      POF_UNROLL_N(UF)
      for (int j = i + 1; j < Cc; ++j) X[r][j] = fma(-w, v[j], X[r][j]);
    }
  }
}

// binomial coefficient at compile time (Pascal blocks of the IWP transition matrix)
POF_HD constexpr double binom(int n, int k) {
  double r = 1.0;
  for (int i = 1; i <= k; ++i) r = r * (double)(n - k + i) / (double)i;
  return r;
}

}  // namespace pof
