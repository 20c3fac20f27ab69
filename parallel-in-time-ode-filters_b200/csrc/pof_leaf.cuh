// Leaf level of the blocked parallel-in-time IEKS: what ONE thread does for ONE chunk of consecutive time
// steps.  Three sequential recursions, all thread-private (registers / local memory), templated on the ODE
// dimension d and the IWP order q (state dimension D = d (q+1), ordering [y1, y1', .., y1^(q), y2, ..],
// reference pof/transitions.py:80-88):
//
//   FoldState   filter phase 1: fold the chunk's leaves into ONE filtering element (A, b, U, eta, Z) of the
//               reference's associative operator (pof/parallel_filtsmooth/filter.py:50-81 builds a leaf,
//               :117-142 combines two).  Folding a *leaf* onto an accumulated element is a conditional
//               square-root Kalman step, ~10x cheaper than the general combine.
//   ScanState   filter phase 3: square-root Kalman filter seeded with the filtered state at the chunk start
//               (the carry from the tree scan).  Per step it yields the innovation statistics (nll, sigma^2;
//               filter.py:84-114), the backward-kernel (g, E, D) of the smoother for that step
//               (smoother.py:37-50; obtained from the SAME QR as the prediction) and it composes the chunk's
//               smoothing element (smoother.py:53-63).
//   SmoothState smoother phase 3: square-root RTS recursion seeded with the smoothed state at the chunk end,
//               objective value (pof/utils.py:97-101 with the reference's swapped arguments,
//               smoother.py:20) and the means-convergence test (pof/convergence_criteria.py:9).
//
// The transition model is the preconditioned IWP (pof/transitions.py:37-50): F = I_d (x) flip(pascal), known at
// compile time (binomials, so F*X is additions only); QL = I_d (x) qL with qL the (q+1)x(q+1) lower Cholesky factor
// of the flipped Hilbert matrix, passed in (constant memory on the device).
#pragma once
#include "pof_small.cuh"

namespace pof {

template <int d, int q>
struct Leaf {
  static constexpr int Q1 = q + 1;
  static constexpr int D = d * Q1;
  static constexpr int UF = 1;  // rolled loops: arrays live in local memory (bring-up / reference path)
  static constexpr double LOG_2PI = 1.8378770664093454835606594728112;

  // M <- F M for an (D x NC) array: within each block, row i <- sum_{j>=i} binom(q-i, j-i) row j.
  template <int NC>
  static POF_HD void mulF(double (&M)[D][NC]) {
    POF_UNROLL_N(UF)
    for (int b = 0; b < d; ++b) {
      POF_UNROLL_N(UF)
      for (int i = 0; i < Q1; ++i) {
        POF_UNROLL_N(UF)
        for (int j = i + 1; j < Q1; ++j) {
          const double cf = binom(q - i, j - i);
          POF_UNROLL_N(UF)
          for (int c = 0; c < NC; ++c) M[b * Q1 + i][c] = fma(cf, M[b * Q1 + j][c], M[b * Q1 + i][c]);
        }
      }
    }
  }
  static POF_HD void mulF_vec(double (&m)[D]) {
    POF_UNROLL_N(UF)
    for (int b = 0; b < d; ++b) {
      POF_UNROLL_N(UF)
      for (int i = 0; i < Q1; ++i) {
        POF_UNROLL_N(UF)
        for (int j = i + 1; j < Q1; ++j) m[b * Q1 + i] = fma(binom(q - i, j - i), m[b * Q1 + j], m[b * Q1 + i]);
      }
    }
  }
  // T <- QL (lower triangular, block diagonal)
  static POF_HD void setQL(double (&T)[D][D], const double* qL) {
    POF_UNROLL_N(UF)
    for (int r = 0; r < D; ++r) {
      POF_UNROLL_N(UF)
      for (int c = 0; c < D; ++c) {
        const bool in = (r / Q1 == c / Q1) && (c % Q1 <= r % Q1);
        T[r][c] = in ? qL[(r % Q1) * Q1 + (c % Q1)] : 0.0;
      }
    }
  }
  // r <- QL^{-1} r (block forward substitution)
  static POF_HD void solveQL(double (&r)[D], const double* qL) {
    POF_UNROLL_N(UF)
    for (int b = 0; b < d; ++b) {
      POF_UNROLL_N(UF)
      for (int i = 0; i < Q1; ++i) {
        double s = r[b * Q1 + i];
        POF_UNROLL_N(UF)
        for (int j = 0; j < i; ++j) s = fma(-qL[i * Q1 + j], r[b * Q1 + j], s);
        r[b * Q1 + i] = s / qL[i * Q1 + i];
      }
    }
  }

  // Measurement update shared by fold and scan.  In: T (D x D lower, predicted factor), H (d x D).
  // Out: SL (d x d lower), Kbar (D x d), Uf (D x D factor of the posterior covariance: first d columns zero).
  static POF_HD void update_factor(const double (&T)[D][D], const double (&H)[d][D], double (&SL)[d][d],
                                   double (&Kbar)[D][d], double (&Uf)[D][D]) {
    double X[d + D][D];
    POF_UNROLL_N(UF)
    for (int a = 0; a < d; ++a) {
      POF_UNROLL_N(UF)
      for (int j = 0; j < D; ++j) {
        double s = 0.0;
        POF_UNROLL_N(UF)
        for (int i = j; i < D; ++i) s = fma(H[a][i], T[i][j], s);
        X[a][j] = s;
      }
    }
    POF_UNROLL_N(UF)
    for (int r = 0; r < D; ++r) {
      POF_UNROLL_N(UF)
      for (int j = 0; j < D; ++j) X[d + r][j] = (j <= r) ? T[r][j] : 0.0;
    }
    house_rows<d + D, D, d, UF>(X);
    POF_UNROLL_N(UF)
    for (int a = 0; a < d; ++a) {
      POF_UNROLL_N(UF)
      for (int j = 0; j < d; ++j) SL[a][j] = (j <= a) ? X[a][j] : 0.0;
    }
    POF_UNROLL_N(UF)
    for (int r = 0; r < D; ++r) {
      POF_UNROLL_N(UF)
      for (int j = 0; j < D; ++j) {
        if (j < d) {
          Kbar[r][j] = X[d + r][j];
          Uf[r][j] = 0.0;
        } else {
          Uf[r][j] = X[d + r][j];
        }
      }
    }
  }
  // z <- SL^{-1} y (forward substitution)
  static POF_HD void solveSL(const double (&SL)[d][d], const double (&y)[d], double (&z)[d]) {
    POF_UNROLL_N(UF)
    for (int a = 0; a < d; ++a) {
      double s = y[a];
      POF_UNROLL_N(UF)
      for (int j = 0; j < a; ++j) s = fma(-SL[a][j], z[j], s);
      z[a] = s / SL[a][a];
    }
  }

  // ------------------------------------------------------------------ filter phase 1
  struct FoldState {
    double A[D][D], b[D], Uf[D][D], eta[D], Z[D][D];
    POF_HD void init() {
      POF_UNROLL_N(UF)
      for (int r = 0; r < D; ++r) {
        b[r] = 0.0;
        eta[r] = 0.0;
        POF_UNROLL_N(UF)
        for (int c = 0; c < D; ++c) {
          A[r][c] = (r == c) ? 1.0 : 0.0;
          Uf[r][c] = 0.0;
          Z[r][c] = 0.0;
        }
      }
    }
    // fold the leaf (H, c) of the next time step onto the accumulated element
    POF_HD void step(const double (&H)[d][D], const double (&c)[d], const double* qL) {
      // predict: A <- F A, b <- F b, T = tria([F Uf, QL])
      mulF<D>(A);
      mulF_vec(b);
      mulF<D>(Uf);
      double T[D][D];
      setQL(T, qL);
      tpqrt<D, D, 0, UF>(T, Uf, (double(*)[D]) nullptr, (double(*)[D]) nullptr);
      // innovation
      double SL[d][d], Kbar[D][d];
      update_factor(T, H, SL, Kbar, Uf);
      // G = SL^{-1} (H A)   (d x D),  z = SL^{-1} (H b + c)
      double G[d][D], r[d], z[d];
      POF_UNROLL_N(UF)
      for (int a = 0; a < d; ++a) {
        double s = c[a];
        POF_UNROLL_N(UF)
        for (int i = 0; i < D; ++i) s = fma(H[a][i], b[i], s);
        r[a] = s;
        POF_UNROLL_N(UF)
        for (int j = 0; j < D; ++j) {
          double t = 0.0;
          POF_UNROLL_N(UF)
          for (int i = 0; i < D; ++i) t = fma(H[a][i], A[i][j], t);
          G[a][j] = t;
        }
      }
      solveSL(SL, r, z);
      POF_UNROLL_N(UF)
      for (int j = 0; j < D; ++j) {
        POF_UNROLL_N(UF)
        for (int a = 0; a < d; ++a) {
          double s = G[a][j];
          POF_UNROLL_N(UF)
          for (int e = 0; e < a; ++e) s = fma(-SL[a][e], G[e][j], s);
          G[a][j] = s / SL[a][a];
        }
      }
      // A <- A - Kbar G ; b <- b - Kbar z ; eta <- eta - G^T z
      POF_UNROLL_N(UF)
      for (int i = 0; i < D; ++i) {
        POF_UNROLL_N(UF)
        for (int a = 0; a < d; ++a) {
          b[i] = fma(-Kbar[i][a], z[a], b[i]);
          eta[i] = fma(-G[a][i], z[a], eta[i]);
          POF_UNROLL_N(UF)
          for (int j = 0; j < D; ++j) A[i][j] = fma(-Kbar[i][a], G[a][j], A[i][j]);
        }
      }
      // Z <- tria([Z, G^T])
      double Gt[D][d];
      POF_UNROLL_N(UF)
      for (int i = 0; i < D; ++i) {
        POF_UNROLL_N(UF)
        for (int a = 0; a < d; ++a) Gt[i][a] = G[a][i];
      }
      tpqrt<D, d, 0, UF>(Z, Gt, (double(*)[D]) nullptr, (double(*)[d]) nullptr);
    }
  };

  // ------------------------------------------------------------------ filter phase 3
  struct StepOut {  // backward kernel of this step + innovation statistics
    double g[D], E[D][D], Dk[D][D];
    double nll, ssq_ref, ssq_proper;
  };
  struct ScanState {
    double m[D], Uf[D][D];
    // one predict+update; (m, Uf) is the filtered state at time k, becomes the one at k+1
    POF_HD void step(const double (&H)[d][D], const double (&c)[d], const double* qL, StepOut& o) {
      double T[D][D], C[D][D];
      POF_UNROLL_N(UF)
      for (int r = 0; r < D; ++r) {
        POF_UNROLL_N(UF)
        for (int j = 0; j < D; ++j) {
          C[r][j] = Uf[r][j];
          o.E[r][j] = 0.0;
        }
      }
      mulF<D>(C);
      setQL(T, qL);
      // [[F Uf, QL],[Uf, 0]] -> [[T, 0],[Phi21, Phi22~]]  (columns permuted: triangular part first)
      tpqrt<D, D, D, UF>(T, C, o.E, Uf);
      // E = Phi21 T^{-1}
      double inv[D];
      POF_UNROLL_N(UF)
      for (int j = 0; j < D; ++j) inv[j] = 1.0 / T[j][j];
      POF_UNROLL_N(UF)
      for (int r = 0; r < D; ++r) {
        POF_UNROLL_N(UF)
        for (int j = D - 1; j >= 0; --j) {
          double s = o.E[r][j];
          POF_UNROLL_N(UF)
          for (int i = j + 1; i < D; ++i) s = fma(-o.E[r][i], T[i][j], s);
          o.E[r][j] = s * inv[j];
        }
      }
      // g = m - E F m ; m <- F m
      double mp[D];
      POF_UNROLL_N(UF)
      for (int i = 0; i < D; ++i) mp[i] = m[i];
      mulF_vec(mp);
      POF_UNROLL_N(UF)
      for (int r = 0; r < D; ++r) {
        double s = m[r];
        POF_UNROLL_N(UF)
        for (int i = 0; i < D; ++i) s = fma(-o.E[r][i], mp[i], s);
        o.g[r] = s;
      }
      // Dk = tria(Phi22~)
      house_rows<D, D, D, UF>(Uf);
      POF_UNROLL_N(UF)
      for (int r = 0; r < D; ++r) {
        POF_UNROLL_N(UF)
        for (int j = 0; j < D; ++j) o.Dk[r][j] = Uf[r][j];
      }
      // update
      double SL[d][d], Kbar[D][d], y[d], z[d];
      update_factor(T, H, SL, Kbar, Uf);
      POF_UNROLL_N(UF)
      for (int a = 0; a < d; ++a) {
        double s = c[a];
        POF_UNROLL_N(UF)
        for (int i = 0; i < D; ++i) s = fma(H[a][i], mp[i], s);
        y[a] = s;
      }
      solveSL(SL, y, z);
      POF_UNROLL_N(UF)
      for (int i = 0; i < D; ++i) {
        double s = mp[i];
        POF_UNROLL_N(UF)
        for (int a = 0; a < d; ++a) s = fma(-Kbar[i][a], z[a], s);
        m[i] = s;
      }
      // innovation statistics: nll = -log N(y; 0, SL SL^T) (pof/utils.py:22-30), ssq_ref = |SL^{-T} y|^2
      // (pof/utils.py:110-112 solves with the transpose), ssq_proper = |SL^{-1} y|^2
      double zz = 0.0, lg = 0.0;
      POF_UNROLL_N(UF)
      for (int a = 0; a < d; ++a) {
        zz = fma(z[a], z[a], zz);
        lg += log(fabs(SL[a][a]));
      }
      o.nll = 0.5 * zz + lg + 0.5 * d * LOG_2PI;
      o.ssq_proper = zz;
      double w[d], ww = 0.0;
      POF_UNROLL_N(UF)
      for (int a = d - 1; a >= 0; --a) {
        double s = y[a];
        POF_UNROLL_N(UF)
        for (int e = a + 1; e < d; ++e) s = fma(-SL[e][a], w[e], s);
        w[a] = s / SL[a][a];
        ww = fma(w[a], w[a], ww);
      }
      o.ssq_ref = ww;
    }
  };

  // smoothing element (g, E, Dm) composition, acc = earlier-in-time, k = later (smoother.py:53-63 with
  // elem1 = later, elem2 = earlier):  g = E_acc g_k + g_acc ; E = E_acc E_k ; Dm = tria([E_acc D_k, D_acc])
  struct SmoothElem {
    double g[D], E[D][D], Dm[D][D];
    POF_HD void set(const StepOut& o) {
      POF_UNROLL_N(UF)
      for (int r = 0; r < D; ++r) {
        g[r] = o.g[r];
        POF_UNROLL_N(UF)
        for (int j = 0; j < D; ++j) {
          E[r][j] = o.E[r][j];
          Dm[r][j] = o.Dk[r][j];
        }
      }
    }
    POF_HD void compose_later(const StepOut& o) {
      double C[D][D];
      POF_UNROLL_N(UF)
      for (int r = 0; r < D; ++r) {
        double s = g[r];
        POF_UNROLL_N(UF)
        for (int i = 0; i < D; ++i) s = fma(E[r][i], o.g[i], s);
        g[r] = s;
        POF_UNROLL_N(UF)
        for (int j = 0; j < D; ++j) {
          double t = 0.0;
          POF_UNROLL_N(UF)
          for (int i = j; i < D; ++i) t = fma(E[r][i], o.Dk[i][j], t);  // D_k lower triangular
          C[r][j] = t;
        }
      }
      POF_UNROLL_N(UF)
      for (int r = 0; r < D; ++r) {
        double row[D];
        POF_UNROLL_N(UF)
        for (int j = 0; j < D; ++j) {
          double t = 0.0;
          POF_UNROLL_N(UF)
          for (int i = 0; i < D; ++i) t = fma(E[r][i], o.E[i][j], t);
          row[j] = t;
        }
        POF_UNROLL_N(UF)
        for (int j = 0; j < D; ++j) E[r][j] = row[j];
      }
      tpqrt<D, D, 0, UF>(Dm, C, (double(*)[D]) nullptr, (double(*)[D]) nullptr);
    }
  };

  // ------------------------------------------------------------------ smoother phase 3
  struct SmoothState {
    double m[D], L[D][D];  // smoothed state at time k+1 (L lower triangular), becomes the one at k
    // returns the objective increment |QL^{-1}(m_k - F m_{k+1})|^2 (reference's swapped-argument form)
    POF_HD double step(const double (&g)[D], const double (&E)[D][D], double (&Dk)[D][D], const double* qL) {
      double C[D][D], mn[D], r[D];
      POF_UNROLL_N(UF)
      for (int i = 0; i < D; ++i) r[i] = m[i];
      mulF_vec(r);
      POF_UNROLL_N(UF)
      for (int a = 0; a < D; ++a) {
        double s = g[a];
        POF_UNROLL_N(UF)
        for (int i = 0; i < D; ++i) s = fma(E[a][i], m[i], s);
        mn[a] = s;
        POF_UNROLL_N(UF)
        for (int j = 0; j < D; ++j) {
          double t = 0.0;
          POF_UNROLL_N(UF)
          for (int i = j; i < D; ++i) t = fma(E[a][i], L[i][j], t);
          C[a][j] = t;
        }
      }
      tpqrt<D, D, 0, UF>(Dk, C, (double(*)[D]) nullptr, (double(*)[D]) nullptr);
      double obj = 0.0;
      POF_UNROLL_N(UF)
      for (int i = 0; i < D; ++i) r[i] = mn[i] - r[i];
      solveQL(r, qL);
      POF_UNROLL_N(UF)
      for (int i = 0; i < D; ++i) {
        obj = fma(r[i], r[i], obj);
        m[i] = mn[i];
        POF_UNROLL_N(UF)
        for (int j = 0; j < D; ++j) L[i][j] = (j <= i) ? Dk[i][j] : 0.0;
      }
      return obj;
    }
  };
};

}  // namespace pof
