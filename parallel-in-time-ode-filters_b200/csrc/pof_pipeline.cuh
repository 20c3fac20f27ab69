// Per-chunk drivers of the blocked parallel-in-time IEKS pass: what one thread (leaf phases) or one warp (tree
// phases) executes.  Shared verbatim by the CUDA kernels (pof_kernels.cu) and the host simulator
// (tests/hostsim) so that the exact device code path is validated against the CPU oracle.
//
// Time indexing: states t = 0..n (N = n+1 grid points), transition/observation k = 0..n-1 takes state k to k+1 and
// uses H[k], c[k] (the linearisation at the previous trajectory's state k+1, reference pof/step.py:12-22).
// Chunk ch owns steps k in [ch*L, min((ch+1)*L, n)).
//
// Per-step backward kernels are kept in a chunk-interleaved struct-of-arrays layout
//     kern[(j * NE + e) * CS + ch],   j = step within chunk, e = element index in [g | E | Dk], CS = #chunks
// so that the 32 threads of a warp (32 consecutive chunks) access consecutive addresses.
#pragma once
#include "pof_coop.cuh"
#include "pof_ivp.cuh"
#include "pof_leaf.cuh"
#include "pof_tree_levels.cuh"

namespace pof {

template <int d, int q>
struct Chunk {
  using LF = Leaf<d, q>;
  static constexpr int D = LF::D;
  static constexpr int UF = LF::UF;
  static constexpr int NE = D + 2 * D * D;  // doubles per step kernel (g, E, Dk)

  static POF_HD void load_Hc(const double* __restrict__ H, const double* __restrict__ c, long k, double (&Hk)[d][D],
                             double (&ck)[d]) {
    POF_UNROLL_N(UF)
    for (int a = 0; a < d; ++a) {
      ck[a] = c[k * d + a];
      POF_UNROLL_N(UF)
      for (int j = 0; j < D; ++j) Hk[a][j] = H[(k * d + a) * D + j];
    }
  }

  // ---- filter phase 1: chunk -> packed filter element
  static POF_HD void fold(long k0, long k1, const double* __restrict__ H, const double* __restrict__ c,
                          const double* qL, double* __restrict__ agg) {
    typename LF::FoldState st;
    st.init();
    for (long k = k0; k < k1; ++k) {
      double Hk[d][D], ck[d];
      load_Hc(H, c, k, Hk, ck);
      st.step(Hk, ck, qL);
    }
    const int DD = D * D;
    POF_UNROLL_N(UF)
    for (int r = 0; r < D; ++r) {
      agg[DD + r] = st.b[r];
      agg[2 * DD + D + r] = st.eta[r];
      POF_UNROLL_N(UF)
      for (int j = 0; j < D; ++j) {
        agg[r * D + j] = st.A[r][j];
        agg[DD + D + r * D + j] = st.Uf[r][j];
        agg[2 * DD + 2 * D + r * D + j] = (j <= r) ? st.Z[r][j] : 0.0;
      }
    }
  }

  // ---- filter phase 3: seeded square-root Kalman filter over the chunk
  //  state_in : packed state (m, L) at the chunk start         sagg : packed smoother element of the chunk (out)
  //  state_end: packed filtered state at the chunk end (out)   part : [nll, ssq_ref, ssq_proper] partial sums (out)
  //  fmeans/fchols (optional, may be null): filtered states k+1 in API layout (N,D),(N,D,D)
  static POF_HD void scan(long k0, long k1, const double* __restrict__ H, const double* __restrict__ c,
                          const double* qL, const double* __restrict__ state_in, double* __restrict__ kern, long CS,
                          long ch, double* __restrict__ sagg, double* __restrict__ state_end,
                          double* __restrict__ part, double* __restrict__ fmeans, double* __restrict__ fchols) {
    typename LF::ScanState st;
    POF_UNROLL_N(UF)
    for (int r = 0; r < D; ++r) {
      st.m[r] = state_in[r];
      POF_UNROLL_N(UF)
      for (int j = 0; j < D; ++j) st.Uf[r][j] = state_in[D + r * D + j];
    }
    typename LF::SmoothElem acc;
    double nll = 0.0, s1 = 0.0, s2 = 0.0;
    for (long k = k0; k < k1; ++k) {
      double Hk[d][D], ck[d];
      load_Hc(H, c, k, Hk, ck);
      typename LF::StepOut o;
      st.step(Hk, ck, qL, o);
      nll += o.nll;
      s1 += o.ssq_ref;
      s2 += o.ssq_proper;
      const long j = k - k0;
      double* kp = kern + (j * NE) * CS + ch;
      POF_UNROLL_N(UF)
      for (int r = 0; r < D; ++r) {
        kp[(long)r * CS] = o.g[r];
        POF_UNROLL_N(UF)
        for (int e = 0; e < D; ++e) {
          kp[(long)(D + r * D + e) * CS] = o.E[r][e];
          kp[(long)(D + D * D + r * D + e) * CS] = o.Dk[r][e];
        }
      }
      if (k == k0) acc.set(o); else acc.compose_later(o);
      if (fmeans) {
        POF_UNROLL_N(UF)
        for (int r = 0; r < D; ++r) {
          fmeans[(k + 1) * D + r] = st.m[r];
          POF_UNROLL_N(UF)
          for (int e = 0; e < D; ++e) fchols[((k + 1) * D + r) * D + e] = st.Uf[r][e];
        }
      }
    }
    // the end state is handed to the smoother tree as a (m, L) state with L lower triangular
    house_rows<D, D, D, UF>(st.Uf);
    POF_UNROLL_N(UF)
    for (int r = 0; r < D; ++r) {
      sagg[r] = acc.g[r];
      state_end[r] = st.m[r];
      POF_UNROLL_N(UF)
      for (int e = 0; e < D; ++e) {
        sagg[D + r * D + e] = acc.E[r][e];
        sagg[D + D * D + r * D + e] = (e <= r) ? acc.Dm[r][e] : 0.0;
        state_end[D + r * D + e] = (e <= r) ? st.Uf[r][e] : 0.0;
      }
    }
    part[0] = nll;
    part[1] = s1;
    part[2] = s2;
  }

  // ---- smoother phase 3: seeded square-root RTS recursion over the chunk (backwards)
  //  seed: packed smoothed state (m, L lower) at time k1.  Writes smoothed states t in [k0, k1) (and t = n when
  //  `last`), scaled by `cscale` (calibration, pof/step.py:42-44).  part: [obj, not_converged_count].
  //  emit_t0: whether this shard owns state row t = 0 (false on ranks > 0 of a time-sharded run)
  static POF_HD void smooth(long k0, long k1, bool last, bool emit_t0, const double* qL,
                            const double* __restrict__ seed,
                            const double* __restrict__ kern, long CS, long ch, double cscale,
                            double* __restrict__ means, double* __restrict__ chols, double* __restrict__ part) {
    typename LF::SmoothState st;
    POF_UNROLL_N(UF)
    for (int r = 0; r < D; ++r) {
      st.m[r] = seed[r];
      POF_UNROLL_N(UF)
      for (int j = 0; j < D; ++j) st.L[r][j] = (j <= r) ? seed[D + r * D + j] : 0.0;
    }
    double obj = 0.0;
    double nconv = 0.0;
    if (last) nconv += emit(k1, st, cscale, means, chols);
    for (long k = k1 - 1; k >= k0; --k) {
      const long j = k - k0;
      const double* kp = kern + (j * NE) * CS + ch;
      double g[D], E[D][D], Dk[D][D];
      POF_UNROLL_N(UF)
      for (int r = 0; r < D; ++r) {
        g[r] = kp[(long)r * CS];
        POF_UNROLL_N(UF)
        for (int e = 0; e < D; ++e) {
          E[r][e] = kp[(long)(D + r * D + e) * CS];
          Dk[r][e] = kp[(long)(D + D * D + r * D + e) * CS];
        }
      }
      obj += st.step(g, E, Dk, qL);
      if (k > 0 || emit_t0) nconv += emit(k, st, cscale, means, chols);
    }
    part[0] = obj;
    part[1] = nconv;
  }

  // ---- sequential EKS: extended Kalman filter relinearised at the PREDICTED mean of every step, then RTS
  //      (reference pof/sequential_filtsmooth/__init__.py:5-10, filter.py:9-30, smoother.py:8-28).  kern: (n, NE)
  //      time-major scratch.  part: [nll, ssq_ref sum, ssq_proper sum, obj].
  static POF_HD void seq_eks(long n, double s0, double s1, const double* qL, int ivp_id, const IvpParams& P,
                             const double* __restrict__ x0, double* __restrict__ kern, double* __restrict__ means,
                             double* __restrict__ chols, double* __restrict__ part) {
    constexpr int Q1 = q + 1;
    typename LF::ScanState st;
    for (int r = 0; r < D; ++r) {
      st.m[r] = x0[r];
      for (int j = 0; j < D; ++j) st.Uf[r][j] = x0[D + r * D + j];
    }
    double nll = 0.0, a1 = 0.0, a2 = 0.0;
    for (long k = 0; k < n; ++k) {
      double mp[D];
      for (int i = 0; i < D; ++i) mp[i] = st.m[i];
      LF::mulF_vec(mp);
      double y[4], f[4], J[16], Hk[d][D], ck[d];
      for (int b = 0; b < d; ++b) y[b] = s0 * mp[b * Q1];
      ivp_eval(ivp_id, P, y, f, J);
      for (int e = 0; e < d; ++e) {
        double ce = -f[e];
        for (int j = 0; j < D; ++j) Hk[e][j] = 0.0;
        for (int b = 0; b < d; ++b) {
          ce = fma(J[e * d + b], y[b], ce);
          Hk[e][b * Q1] = -J[e * d + b] * s0;
        }
        Hk[e][e * Q1 + 1] += s1;
        ck[e] = ce;
      }
      typename LF::StepOut o;
      st.step(Hk, ck, qL, o);
      nll += o.nll;
      a1 += o.ssq_ref;
      a2 += o.ssq_proper;
      double* kp = kern + k * NE;
      for (int r = 0; r < D; ++r) {
        kp[r] = o.g[r];
        for (int e = 0; e < D; ++e) {
          kp[D + r * D + e] = o.E[r][e];
          kp[D + D * D + r * D + e] = o.Dk[r][e];
        }
      }
    }
    house_rows<D, D, D, 1>(st.Uf);
    typename LF::SmoothState sm;
    for (int r = 0; r < D; ++r) {
      sm.m[r] = st.m[r];
      for (int j = 0; j < D; ++j) sm.L[r][j] = (j <= r) ? st.Uf[r][j] : 0.0;
    }
    double obj = 0.0;
    emit(n, sm, 1.0, means, chols);
    for (long k = n - 1; k >= 0; --k) {
      const double* kp = kern + k * NE;
      double g[D], E[D][D], Dk[D][D];
      for (int r = 0; r < D; ++r) {
        g[r] = kp[r];
        for (int e = 0; e < D; ++e) {
          E[r][e] = kp[D + r * D + e];
          Dk[r][e] = kp[D + D * D + r * D + e];
        }
      }
      obj += sm.step(g, E, Dk, qL);
      emit(k, sm, 1.0, means, chols);
    }
    part[0] = nll;
    part[1] = a1;
    part[2] = a2;
    part[3] = obj;
  }

  // write one smoothed state in API layout; returns the number of mean entries that fail
  // isclose(old, new, rtol=1e-13, atol=1e-8)  (pof/convergence_criteria.py:9; numpy/jax isclose semantics)
  static POF_HD double emit(long t, const typename LF::SmoothState& st, double cscale, double* __restrict__ means,
                            double* __restrict__ chols) {
    double bad = 0.0;
    POF_UNROLL_N(UF)
    for (int r = 0; r < D; ++r) {
      const double old = means[t * D + r];
      const double nw = st.m[r];
      // isclose(a=old, b=new): |a-b| <= atol + rtol*|b| ; NaN never close
      const bool close = fabs(old - nw) <= (1e-8 + 1e-13 * fabs(nw));
      bad += close ? 0.0 : 1.0;
      means[t * D + r] = nw;
      if (chols) {
        POF_UNROLL_N(UF)
        for (int e = 0; e < D; ++e) chols[(t * D + r) * D + e] = (e <= r) ? cscale * st.L[r][e] : 0.0;
      }
    }
    return bad;
  }
};

}  // namespace pof
