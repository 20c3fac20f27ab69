// Large-state path of the blocked parallel-in-time IEKS ("tile" family): runtime ODE dimension d and IWP order q,
// state dimension D = d (q+1) up to ~64 (BASELINE config 5: Lorenz-96, d = 16, q = 3, D = 64).  One CTA works on one
// time-chunk (leaf recursions) or on one tree node (associative combines); the matrices live in shared memory
// tiles, the rows of every Householder triangularisation are spread over the threads of the CTA.
//
// The recursions are the same as in the register-resident families (pof_leaf.cuh states the math; reference
// pof/parallel_filtsmooth/filter.py:50-142, smoother.py:37-63, sequential_filtsmooth/filter.py:60-92):
//   tile_fold    filter phase 1: fold a chunk's leaves into ONE filtering element (A, b, U, eta, Z) and the same
//                element taken before the chunk's last measurement update
//   tile_scan    filter phase 3: square-root Kalman filter seeded with the chunk's incoming filtered state; per step
//                the innovation statistics (nll, sigma^2) and the backward kernel (g, E, Dk) of the smoother
//   tile_smooth  smoother phase 3: square-root RTS recursion seeded with the smoothed state at the chunk end
//   tile_filter_combine / tile_smooth_combine / tile_chunk_kernel   the tree operators on general elements
//
// Programming model.  Every piece of work is a `Team::each(n, body)`: the n iterations are independent, they are
// dealt to the threads of the CTA, and a CTA barrier follows.  ALL writes to shared or global memory happen inside
// such bodies; nothing outside a body reads memory.  So the code is race-free iff the iterations of each single
// `each` are independent of each other -- which the host simulator (tests/hostsim) checks by executing them in
// forward, reverse and shuffled order (the results must be bitwise identical) besides checking the math against the
// oracle.  Control flow around `each` only depends on CTA-uniform values.  (`Team::each_group` is the one extension:
// H threads per row exchange partial sums through shuffles between a stage 1 and a stage 2; see there.)
//
// Code shape.  The leaf recursions are loops over PHASES; a phase does its element-wise / GEMM-like work and leaves at
// most one Householder-sweep request, executed at ONE inlined call site at the bottom of the loop.  This is what lets
// ptxas keep the register-resident sweeps in registers (DESIGN.md 2.4 lists what did not work).
//
// Observation noise: `R` (the reference's cholR, observations.py:23-33) may be non-zero here -- the posterior factor
// then has D instead of D-d non-zero columns; R == nullptr is the noiseless ODE-solver case (step.py:12-22).
#pragma once
#include <vector>

#include "pof_ivp.cuh"
#include "pof_small.cuh"

namespace pof {

#if defined(__CUDACC__)
#define POF_TDEV __host__ __device__ __forceinline__
// the Householder sweeps: inlined, one call site per leaf kernel (see "Code shape" above)
#define POF_TFUNC __host__ __device__ __forceinline__
#else
#define POF_TDEV inline
#define POF_TFUNC inline
#endif

struct Team {
#if defined(__CUDA_ARCH__)
  int tid, nt;
  __device__ __forceinline__ Team() : tid(threadIdx.x), nt(blockDim.x) {}
  template <class F>
  __device__ __forceinline__ void each(int n, F&& f) const {
    for (int i = tid; i < n; i += nt) f(i);
    __syncthreads();
  }
  __device__ __forceinline__ int threads() const { return nt; }
  // n rows x H cooperating threads per row (H a power of two <= 32, n * H <= #threads, so thread tid IS (row tid / H,
  // part tid % H) in every call).  stage1(row, part) returns two partial sums; they are added over the H parts of a
  // row (xor butterfly: all H threads get bitwise the same totals); stage2(row, part, totals) finishes.  Barrier.
  // Everything a stage-2 body reads that another part's stage 2 writes must travel through the three sum channels.
  template <int H, class S1, class S2>
  __device__ __forceinline__ void each_group(int n, S1&& s1, S2&& s2) const {
    const bool active = tid < n * H;
    const int r = tid / H, h = tid % H;
    double a = 0.0, b = 0.0, c = 0.0;
    if (active) s1(r, h, a, b, c);
#pragma unroll
    for (int o = 1; o < H; o <<= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      c += __shfl_xor_sync(0xffffffffu, c, o);
    }
    if (active) s2(r, h, a, b, c);
    __syncthreads();
  }
#else
  int threads() const { return 256; }  // the CTA size of the kernels (TILE_THREADS): same dispatch decisions as the device
  template <int H, class S1, class S2>
  void each_group(int n, S1&& s1, S2&& s2) const {
    auto row = [&](int r) {
      double a[H], b[H], c[H];
      for (int h = 0; h < H; ++h) {
        a[h] = b[h] = c[h] = 0.0;
        s1(r, h, a[h], b[h], c[h]);
      }
      for (int o = 1; o < H; o <<= 1) {  // the same butterfly as the shuffles
        double na[H], nb[H], nc[H];
        for (int h = 0; h < H; ++h) {
          na[h] = a[h] + a[h ^ o];
          nb[h] = b[h] + b[h ^ o];
          nc[h] = c[h] + c[h ^ o];
        }
        for (int h = 0; h < H; ++h) {
          a[h] = na[h];
          b[h] = nb[h];
          c[h] = nc[h];
        }
      }
      // stage 2 of one part must not depend on stage 2 of another one: forward / reverse with the iteration order
      if (order() == 0) {
        for (int h = 0; h < H; ++h) s2(r, h, a[h], b[h], c[h]);
      } else {
        for (int h = H - 1; h >= 0; --h) s2(r, h, a[h], b[h], c[h]);
      }
    };
    const int o = order();
    if (o == 0) {
      for (int r = 0; r < n; ++r) row(r);
    } else {
      for (int r = n - 1; r >= 0; --r) row(r);
    }
  }
  // host simulator: one "thread" runs all iterations; the order is selectable to expose cross-iteration dependences
  static int& order() {
    static int o = 0;
    return o;
  }
  template <class F>
  void each(int n, F&& f) const {
    const int o = order();
    if (o == 0) {
      for (int i = 0; i < n; ++i) f(i);
    } else if (o == 1) {
      for (int i = n - 1; i >= 0; --i) f(i);
    } else {  // a fixed pseudo-random permutation: i -> (a i + b) mod n with gcd(a, n) = 1
      int a = 7919 % (n > 0 ? n : 1);
      if (a == 0) a = 1;
      auto gcd = [](int x, int y) { while (y) { int t = x % y; x = y; y = t; } return x; };
      while (gcd(a, n) != 1) ++a;
      for (int i = 0; i < n; ++i) f((int)(((long)a * i + 3) % n));
    }
  }
#endif
};

// init + sum_{k in [lo, hi)} a(k) b(k) with four independent FMA chains (a dependent DFMA issues every ~8 cycles on
// B200, and a CTA of 8 warps cannot hide that by itself)
template <class FA, class FB>
POF_TDEV double tile_dot(int lo, int hi, double init, FA a, FB b) {
  double s0 = init, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int k = lo;
  for (; k + 3 < hi; k += 4) {
    s0 = fma(a(k), b(k), s0);
    s1 = fma(a(k + 1), b(k + 1), s1);
    s2 = fma(a(k + 2), b(k + 2), s2);
    s3 = fma(a(k + 3), b(k + 3), s3);
  }
  for (; k < hi; ++k) s0 = fma(a(k), b(k), s0);
  return (s0 + s1) + (s2 + s3);
}

constexpr int TILE_PB_COLS_ = 128;  // columns of one pivot-row broadcast buffer (register sweeps)

// ---------------------------------------------------------------------------------------------------------------
// Right-Householder lower-triangularisation of M (R x C, leading dimension ld), pivots 0..npiv-1, LAPACK sign
// convention (beta = -sign(alpha) norm), Q never formed (replaces the reference's tria(), pof/utils.py:33-41).
//   c0 <  0 : plain.  Pivot i eliminates columns (i, C) of row i.
//   c0 >= 0 : triangular-pentagonal.  Columns [0, c0) of the pivot rows are lower triangular already (entries right of
//             the diagonal there are never referenced); pivot i eliminates columns [c0, C) of row i.
// One barrier per pivot: thread `rr` > 0 owns row i+rr, re-derives the reflector of pivot row i from shared memory
// (fused with its own dot product, so the pivot row is read once and never rescaled in memory) and updates its row;
// thread 0 only records beta.  The diagonal is written back at the end; the eliminated entries of the pivot rows
// keep stale values (never referenced afterwards: only the lower triangle / the non-pivot rows are meaningful).
// ---------------------------------------------------------------------------------------------------------------
POF_TFUNC void tile_tria_smem(const Team& t, double* M, int R, int C, int ld, int npiv, int c0, double* diag) {
  for (int i = 0; i < npiv; ++i) {
    const int js = c0 >= 0 ? c0 : i + 1;
    t.each(R - i, [&](int rr) {
      const double* p = M + (long)i * ld;
      double* row = M + (long)(i + rr) * ld;
      double s0 = 0.0, s1 = 0.0, d0 = 0.0, d1 = 0.0;
      int j = js;
      for (; j + 1 < C; j += 2) {
        const double p0 = p[j], p1 = p[j + 1];
        s0 = fma(p0, p0, s0);
        s1 = fma(p1, p1, s1);
        d0 = fma(row[j], p0, d0);
        d1 = fma(row[j + 1], p1, d1);
      }
      if (j < C) {
        const double p0 = p[j];
        s0 = fma(p0, p0, s0);
        d0 = fma(row[j], p0, d0);
      }
      const double sigma = s0 + s1;
      const double alpha = p[i];
      if (!(sigma > 0.0)) {  // nothing to eliminate: identity (also keeps all-zero rows free of NaN)
        if (rr == 0) diag[i] = alpha;
        return;
      }
      const double nrm = sqrt(fma(alpha, alpha, sigma));
      const double beta = (alpha >= 0.0) ? -nrm : nrm;
      if (rr == 0) {
        diag[i] = beta;
        return;
      }
      const double tau = (beta - alpha) / beta;
      const double scale = 1.0 / (alpha - beta);
      // v = (1, scale * tail); w = tau * row . v
      const double w = tau * fma(scale, d0 + d1, row[i]);
      row[i] -= w;
      const double ws = w * scale;
      for (j = js; j < C; ++j) row[j] = fma(-ws, p[j], row[j]);
    });
  }
  t.each(npiv, [&](int i) { M[(long)i * ld + i] = diag[i]; });
}

// ---------------------------------------------------------------------------------------------------------------
// The same triangularisation with the rows held in REGISTERS for the whole sweep (the shared-memory version re-reads
// and re-writes the trailing matrix for every pivot: 8 shared-memory wavefronts per column and pivot; here a pivot
// costs one broadcast read of the pivot row per column).  `regs.at(it)` is the executing thread's register file on
// the device (iteration `it` of every each() / each_group() below runs on thread `it`, because the iteration counts
// do not exceed the thread count) and a per-iteration array in the host simulator; it must be called with the
// iteration index of the enclosing each, never with a derived row number.  All loops over the registers are fully
// unrolled so that nothing is indexed dynamically.  The pivot row travels through a double-buffered broadcast array
// pb (2 x TILE_PB_COLS doubles of shared memory): the owners of row i+1 publish it at the end of pivot i, so there is
// still ONE barrier per pivot.  The reflector is H = I - tp v v^T, v = (alpha - beta, tail), tp = 1 / (norm |v_0|),
// derived with the hardware reciprocal / reciprocal-square-root approximations plus two Newton steps, as in the
// register-resident leaf kernels (pof_lane2.cuh): the fp64 division / square-root subroutine calls made ptxas spill
// the rows around them.
// ---------------------------------------------------------------------------------------------------------------
POF_TDEV double tile_fast_rcp(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
#else
  return 1.0 / x;
#endif
}
POF_TDEV double tile_fast_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  const double hx = 0.5 * x;
  double e = fma(-hx * r, r, 0.5);
  r = fma(r, e, r);
  e = fma(-hx * r, r, 0.5);
  return fma(r, e, r);
#else
  return 1.0 / sqrt(x);
#endif
}
// (beta, v0 = alpha - beta, tp) of the reflector for a row with head alpha and squared tail norm sigma; ok = false if
// there is nothing to eliminate (zero tail; the approximations flush subnormals: squared norms below 2^-1000 count as
// zero) -- then the reflector is the identity and beta = alpha
struct TileHouse {
  double beta, v0, tp;
  bool ok;
};
POF_TDEV TileHouse tile_house(double alpha, double sigma) {
  TileHouse h;
  const double nrm2 = fma(alpha, alpha, sigma);
  h.ok = sigma > 0.0 && nrm2 > 0x1p-1000;
  const double rn = tile_fast_rsqrt(h.ok ? nrm2 : 1.0);
  const double nrm = nrm2 * rn;
  h.beta = h.ok ? ((alpha >= 0.0) ? -nrm : nrm) : alpha;
  h.v0 = alpha - h.beta;
  h.tp = h.ok ? rn * tile_fast_rcp(fabs(h.v0)) : 0.0;
  return h;
}

template <int KT>
struct RowRegs {
#if defined(__CUDA_ARCH__)
  double v[KT];
  __device__ __forceinline__ explicit RowRegs(int) {}
  __device__ __forceinline__ double* at(int) { return v; }
#else
  std::vector<double> store;
  explicit RowRegs(int its) : store((size_t)its * KT) {}
  double* at(int it) { return &store[(size_t)it * KT]; }
#endif
};

// Pentagonal mode ([T | C], T lower triangular in shared memory, C = columns [c0, C) in registers), H threads per
// row: thread (r, h) keeps columns [h KT, (h+1) KT) of C.  R H <= #threads, C - c0 <= KT H.
template <int KT, int H>
POF_TFUNC void tile_tria_regp(const Team& t, double* M, int R, int C, int ld, int npiv, int c0, double* diag,
                              double* pb) {
  const int K = C - c0;
  RowRegs<KT> regs(R * H);
  t.each(R * H, [&](int it) {
    const int r = it / H, h = it % H;
    double* rv = regs.at(it);
    const double* row = M + (long)r * ld + c0 + h * KT;
    POF_UNROLL_N(KT)
    for (int j = 0; j < KT; ++j) rv[j] = (h * KT + j < K) ? row[j] : 0.0;
    if (r == 0) {
      POF_UNROLL_N(KT)
      for (int j = 0; j < KT; ++j) pb[h * KT + j] = rv[j];
    }
  });
  for (int i = 0; i < npiv; ++i) {
    const double* p = pb + (i & 1) * TILE_PB_COLS_;
    double* pn = pb + ((i + 1) & 1) * TILE_PB_COLS_;
    t.template each_group<H>(
        R,
        // partial squared norm of the pivot row, partial dot product; part 0 also contributes the row's entry in the
        // pivot column (it is rewritten by part 0 in stage 2, so the other parts must not read it there)
        [&](int r, int h, double& sg, double& dt, double& ri) {
          if (r < i) return;
          if (h == 0) ri = M[(long)r * ld + i];
          const double* rv = regs.at(r * H + h);
          const double* ph = p + h * KT;
          double s[4] = {0.0, 0.0, 0.0, 0.0}, d[4] = {0.0, 0.0, 0.0, 0.0};
          POF_UNROLL_N(KT)
          for (int j = 0; j < KT; ++j) {
            const double pj = ph[j];
            s[j & 3] = fma(pj, pj, s[j & 3]);
            d[j & 3] = fma(rv[j], pj, d[j & 3]);
          }
          sg = (s[0] + s[1]) + (s[2] + s[3]);
          dt = (d[0] + d[1]) + (d[2] + d[3]);
        },
        [&](int r, int h, double sigma, double dot, double rowi) {
          if (r < i) return;
          double* rv = regs.at(r * H + h);
          const double* ph = p + h * KT;
          const TileHouse hh = tile_house(M[(long)i * ld + i], sigma);
          if (r == i) {
            if (h == 0) diag[i] = hh.beta;
            return;
          }
          const double ws = hh.tp * fma(hh.v0, rowi, dot);
          if (h == 0) M[(long)r * ld + i] = fma(-ws, hh.v0, rowi);  // column i is final now
          POF_UNROLL_N(KT)
          for (int j = 0; j < KT; ++j) rv[j] = fma(-ws, ph[j], rv[j]);
          if (r == i + 1) {  // publish the next pivot row
            POF_UNROLL_N(KT)
            for (int j = 0; j < KT; ++j) pn[h * KT + j] = rv[j];
          }
        });
  }
  // columns right of the triangular block: only the non-pivot rows carry meaningful values there
  t.each(R * H, [&](int it) {
    const int r = it / H, h = it % H;
    if (r < npiv) return;
    const double* rv = regs.at(it);
    double* row = M + (long)r * ld + c0 + h * KT;
    POF_UNROLL_N(KT)
    for (int j = 0; j < KT; ++j)
      if (h * KT + j < K) row[j] = rv[j];
  });
  t.each(npiv, [&](int i) { M[(long)i * ld + i] = diag[i]; });
}

// Plain mode, one thread per row: the registers hold columns [i, C) of the CURRENT pivot i -- after every pivot the
// finished column i is written to shared memory and the register file shifts left by one (compile-time indices only:
// a `rv[i]` with a runtime i would send the whole array to local memory).  R <= #threads, C <= KT.
template <int KT>
POF_TFUNC void tile_tria_reg1(const Team& t, double* M, int R, int C, int ld, int npiv, double* diag, double* pb) {
  RowRegs<KT> regs(R);
  t.each(R, [&](int r) {
    double* rv = regs.at(r);
    const double* row = M + (long)r * ld;
    POF_UNROLL_N(KT)
    for (int j = 0; j < KT; ++j) rv[j] = (j < C) ? row[j] : 0.0;
    if (r == 0) {
      POF_UNROLL_N(KT)
      for (int j = 0; j < KT; ++j) pb[j] = rv[j];
    }
  });
  for (int i = 0; i < npiv; ++i) {
    const double* p = pb + (i & 1) * TILE_PB_COLS_;
    double* pn = pb + ((i + 1) & 1) * TILE_PB_COLS_;
    // all R iterations, not R - i: iteration r must stay on thread r, whose registers hold row r
    t.each(R, [&](int r) {
      if (r < i) return;
      double* rv = regs.at(r);
      double s[4] = {0.0, 0.0, 0.0, 0.0}, d[4] = {0.0, 0.0, 0.0, 0.0};
      POF_UNROLL_N(KT)
      for (int j = 1; j < KT; ++j) {  // columns beyond C hold zeros
        const double pj = p[j];
        s[j & 3] = fma(pj, pj, s[j & 3]);
        d[j & 3] = fma(rv[j], pj, d[j & 3]);
      }
      const TileHouse hh = tile_house(p[0], (s[0] + s[1]) + (s[2] + s[3]));
      if (r == i) {
        diag[i] = hh.beta;
        return;
      }
      const double rowi = rv[0];
      const double ws = hh.tp * fma(hh.v0, rowi, (d[0] + d[1]) + (d[2] + d[3]));
      M[(long)r * ld + i] = fma(-ws, hh.v0, rowi);  // column i is final now
      POF_UNROLL_N(KT)
      for (int j = 1; j < KT; ++j) rv[j - 1] = fma(-ws, p[j], rv[j]);  // update and shift left by one
      rv[KT - 1] = 0.0;
      if (r == i + 1) {  // publish the next pivot row
        POF_UNROLL_N(KT)
        for (int j = 0; j < KT; ++j) pn[j] = rv[j];
      }
    });
  }
  t.each(R, [&](int r) {
    if (r < npiv) return;
    const double* rv = regs.at(r);
    double* row = M + (long)r * ld + npiv;
    POF_UNROLL_N(KT)
    for (int j = 0; j < KT; ++j)
      if (npiv + j < C) row[j] = rv[j];
  });
  t.each(npiv, [&](int i) { M[(long)i * ld + i] = diag[i]; });
}

// pb: 2 x TILE_PB_COLS doubles of shared memory for the register sweeps, or null to force the shared-memory sweep.
// Pentagonal sweeps: two threads per row, up to 32 columns per thread; plain sweeps: one thread per row, up to 32
// columns.  Everything else (the 2D x 2D arrays of the tree operators and the plain sweeps
// at D = 64) runs the shared-memory sweep.
POF_TDEV void tile_tria(const Team& t, double* M, int R, int C, int ld, int npiv, int c0, double* diag, double* pb) {
  const int nt = t.threads();
  if (pb != nullptr && c0 >= 0) {
    // two threads per row, 8 / 16 / 32 columns each (few instantiations on purpose: with six of them inlined into one
    // kernel ptxas kept the register files of some in local memory)
    const int K = C - c0;
    if (R * 2 <= nt && K <= 64) {
      if (K <= 16)
        tile_tria_regp<8, 2>(t, M, R, C, ld, npiv, c0, diag, pb);
      else if (K <= 32)
        tile_tria_regp<16, 2>(t, M, R, C, ld, npiv, c0, diag, pb);
      else
        tile_tria_regp<32, 2>(t, M, R, C, ld, npiv, c0, diag, pb);
      return;
    }
  } else if (pb != nullptr && R <= nt && C <= 32) {
    if (C <= 16)
      tile_tria_reg1<16>(t, M, R, C, ld, npiv, diag, pb);
    else
      tile_tria_reg1<32>(t, M, R, C, ld, npiv, diag, pb);
    return;
  }
  tile_tria_smem(t, M, R, C, ld, npiv, c0, diag);
}

// ---------------------------------------------------------------------------------------------------------------
// model constants held in shared memory: binomials of the IWP transition matrix and the (q+1)x(q+1) block of QL
// ---------------------------------------------------------------------------------------------------------------
struct TileModel {
  int d, q, Q1, D;
  const double* cf;  // Q1 x Q1: cf[i][j] = binom(q-i, j-i) for j >= i (row i of flip(pascal)), else 0
  const double* ql;  // Q1 x Q1 lower triangular
  // general transition model of the current step (the reference's per-step TransitionModel(F, QL), e.g. the
  // non-preconditioned / non-uniform-grid models of pof/transitions.py:71-99): dense D x D in global memory, or null =
  // the preconditioned IWP above
  const double* Fd;
  const double* Qd;
};
constexpr int TILE_MODEL_DOUBLES = 2 * 36;

POF_TDEV void tile_model_init(const Team& t, TileModel& md, int d, int q, const double* ql_param, double* smem) {
  md.d = d;
  md.q = q;
  md.Q1 = q + 1;
  md.D = d * (q + 1);
  double* cf = smem;
  double* ql = smem + 36;
  md.cf = cf;
  md.ql = ql;
  md.Fd = nullptr;
  md.Qd = nullptr;
  const int Q1 = q + 1;
  t.each(Q1 * Q1, [&](int idx) {
    const int i = idx / Q1, j = idx - i * Q1;
    double r = 0.0;
    if (j >= i) {
      r = 1.0;
      const int n = q - i, k = j - i;
      for (int u = 1; u <= k; ++u) r = r * (double)(n - k + u) / (double)u;
    }
    cf[idx] = r;
    ql[idx] = ql_param[idx];
  });
}
// (F X)(r, c) for X given by an accessor: row r = b*Q1 + i  <-  sum_{j >= i} cf[i][j] X(b*Q1 + j, c)
template <class FX>
POF_TDEV double tile_F_row(const TileModel& md, int r, FX x) {
  double s = 0.0;
  if (md.Fd) {
    s = tile_dot(0, md.D, s, [&](int j) { return md.Fd[r * md.D + j]; }, [&](int j) { return x(j); });
    return s;
  }
  const int b = r / md.Q1, i = r - b * md.Q1;
  s = tile_dot(i, md.Q1, s, [&](int j) { return md.cf[i * md.Q1 + j]; }, [&](int j) { return x(b * md.Q1 + j); });
  return s;
}
POF_TDEV double tile_QL(const TileModel& md, int r, int c) {
  if (md.Qd) return (c <= r) ? md.Qd[r * md.D + c] : 0.0;
  const int br = r / md.Q1, bc = c / md.Q1;
  const int i = r - br * md.Q1, j = c - bc * md.Q1;
  return (br == bc && j <= i) ? md.ql[i * md.Q1 + j] : 0.0;
}

// the linearisation of one step staged in shared memory: Hs (d x D), cs (d), Rs (d x d, lower) -- from the dense
// arrays (H, c) or from the compact form [J_f | c] with H = E1 - J_f E0 (E0 = s0 e_0^T, E1 = s1 e_1^T per block)
struct TileLin {
  const double* H;   // (n, d, D) or null
  const double* c;   // (n, d)
  const double* Jc;  // (n, d*d + d) or null
  const double* R;   // (n, d, d) or null (noiseless)
  double s0, s1;
  const double* F;   // general per-step transition matrices (n, D, D) and noise factors (lower), or null = IWP
  const double* QL;
  int reg_sweeps;    // != 0: Householder sweeps with register-resident rows where they apply (tile_tria), else the
                     // shared-memory sweep everywhere (default until the register sweeps have been timed on a B200)
};
POF_TDEV void tile_set_step_model(TileModel& md, const TileLin& lin, long k) {
  md.Fd = lin.F ? lin.F + k * md.D * md.D : nullptr;
  md.Qd = lin.QL ? lin.QL + k * md.D * md.D : nullptr;
}
// Sequential EKS (reference pof/sequential_filtsmooth/filter.py:9-30): the observation model is relinearised at the
// PREDICTED mean of every step, inside the kernel, for a built-in vector field.
struct TileEks {
  int ivp_id;
  IvpParams P;
};
// Hs, cs <- linearisation of x -> E1 x - f(E0 x) at the predicted mean mp; Rs <- 0
POF_TDEV void tile_stage_lin_eks(const Team& t, const TileModel& md, const TileEks& eks, double s0, double s1,
                                 const double* mp, double* Hs, double* cs, double* Rs) {
  const int d = md.d, D = md.D, Q1 = md.Q1;
  t.each(d, [&](int a) {
    for (int e = 0; e < d; ++e) Rs[a * d + e] = 0.0;
    if (eks.ivp_id == POF_IVP_LORENZ96) {
      l96_linearize_row(eks.P.p[0], 0, a, d, md.q, s0, s1, 1, mp, Hs, cs, nullptr);
    } else if (a == 0) {  // d <= 4: one iteration evaluates the whole field and Jacobian
      double y[4], f[4], J[16];
      for (int b = 0; b < d; ++b) y[b] = s0 * mp[b * Q1];
      ivp_eval(eks.ivp_id, eks.P, y, f, J);
      for (int r = 0; r < d; ++r) {
        double ce = -f[r];
        for (int j = 0; j < D; ++j) Hs[r * D + j] = 0.0;
        for (int b = 0; b < d; ++b) {
          ce = fma(J[r * d + b], y[b], ce);
          Hs[r * D + b * Q1] = -J[r * d + b] * s0;
        }
        Hs[r * D + r * Q1 + 1] += s1;
        cs[r] = ce;
      }
    }
  });
}
POF_TDEV void tile_stage_lin(const Team& t, const TileModel& md, const TileLin& lin, long k, double* Hs, double* cs,
                             double* Rs) {
  const int d = md.d, D = md.D, Q1 = md.Q1;
  t.each(d * (D + 1 + d), [&](int idx) {
    const int a = idx / (D + 1 + d), j = idx - a * (D + 1 + d);
    if (j < D) {
      double h;
      if (lin.Jc) {
        const int b = j / Q1, i = j - b * Q1;
        h = 0.0;
        if (i == 0) h = -lin.Jc[k * (d * d + d) + a * d + b] * lin.s0;
        if (i == 1 && b == a) h += lin.s1;
      } else {
        h = lin.H[(k * d + a) * D + j];
      }
      Hs[a * D + j] = h;
    } else if (j == D) {
      cs[a] = lin.Jc ? lin.Jc[k * (d * d + d) + d * d + a] : lin.c[k * d + a];
    } else {
      const int e = j - D - 1;
      Rs[a * d + e] = lin.R ? lin.R[(k * d + a) * d + e] : 0.0;
    }
  });
}

// ---------------------------------------------------------------------------------------------------------------
// Measurement update shared by fold and scan.  In: the predicted factor T (lower, D x D, leading dimension ldT) and
// the staged linearisation.  X ((d+D) x (D+d), leading dimension ldx = D+d+1) is built as [[H T, R],[T, 0]] and its
// first d rows are triangularised.  Afterwards  SL = X[0:d, 0:d] (lower),  Kbar = X[d:, 0:d],  posterior factor
// Uf = X[d:, d:d+D] (its last d columns are zero when R == 0).
// ---------------------------------------------------------------------------------------------------------------
POF_TDEV void tile_update_build(const Team& t, const TileModel& md, const double* T, int ldT, const double* Hs,
                                const double* Rs, double* X) {
  const int d = md.d, D = md.D, ldx = D + d + 1, W = D + d;
  t.each((d + D) * W, [&](int idx) {
    const int r = idx / W, j = idx - r * W;
    double v = 0.0;
    if (r < d) {
      if (j < D) {
        v = tile_dot(j, D, v, [&](int i) { return Hs[r * D + i]; }, [&](int i) { return T[(long)i * ldT + j]; });
      } else {
        v = Rs[r * d + (j - D)];
      }
    } else if (j < D) {
      v = (j <= r - d) ? T[(long)(r - d) * ldT + j] : 0.0;
    }
    X[(long)r * ldx + j] = v;
  });
}
// One Householder sweep request.  The leaf recursions below are written as loops over PHASES: every phase does its
// element-wise / GEMM-like work and may leave one sweep request, which is executed at the single tile_tria call site at
// the bottom of the loop -- so each kernel contains ONE inlined copy of the sweeps (register allocation is then global
// and the sweeps are not squeezed into whatever registers a caller leaves free across a call).
struct TileSweep {
  double* M;
  int R, C, ld, npiv, c0;
  POF_TDEV void none() { M = nullptr; }
  POF_TDEV void set(double* M_, int R_, int C_, int ld_, int npiv_, int c0_) {
    M = M_;
    R = R_;
    C = C_;
    ld = ld_;
    npiv = npiv_;
    c0 = c0_;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// shared-memory layouts (in doubles) of the three leaf kernels
// ---------------------------------------------------------------------------------------------------------------
POF_TDEV int tile_vec_doubles(int D, int d) { return 8 * D + 4 * d + d * d + 16 + 2 * TILE_PB_COLS_; }
POF_TDEV int tile_fold_smem_doubles(int D, int d) {
  return TILE_MODEL_DOUBLES + D * (2 * D + 1) + (D + d) * (D + d + 1) + D * (D + 1) + D * (D + d + 1) +
         d * (D + 1) + d * D + tile_vec_doubles(D, d);
}
POF_TDEV int tile_scan_smem_doubles(int D, int d) {
  return TILE_MODEL_DOUBLES + 2 * D * (2 * D + 1) + (D + d) * (D + d + 1) + d * D + tile_vec_doubles(D, d);
}
POF_TDEV int tile_smooth_smem_doubles(int D, int d) {
  return TILE_MODEL_DOUBLES + D * (2 * D + 1) + 2 * D * (D + 1) + tile_vec_doubles(D, d);
}
POF_TDEV int tile_tree_smem_doubles(int D) { return 2 * D * (2 * D + 1) + D * (2 * D + 1) + 10 * D + 16; }

// small vectors common to the leaf kernels
struct TileVecs {
  double *v0, *v1, *v2, *v3;  // D each
  double *diag;               // 2D  (+2D spare)
  double *cs, *y, *z, *w;     // d each
  double *Rs;                 // d x d
  double *acc;                // 16 scalars
  double *pb;                 // 2 x TILE_PB_COLS: pivot-row broadcast buffers of tile_tria_reg
  POF_TDEV void init(double* base, int D, int d) {
    v0 = base;
    v1 = v0 + D;
    v2 = v1 + D;
    v3 = v2 + D;
    diag = v3 + D;
    cs = diag + 4 * D;
    y = cs + d;
    z = y + d;
    w = z + d;
    Rs = w + d;
    acc = Rs + d * d;
    pb = acc + 16;
  }
};

// ---------------------------------------------------------------------------------------------------------------
// filter phase 1
// ---------------------------------------------------------------------------------------------------------------
// agg / aggm: packed filtering elements [A D*D | b D | U D*D | eta D | Z D*D]; aggm (may be null) is the element before
// the chunk's last measurement update (consumed by tile_chunk_kernel).
POF_TDEV void tile_fold(const Team& t, int d, int q, const double* ql_param, const TileLin& lin, long k0, long k1,
                        double* __restrict__ agg, double* __restrict__ aggm, double* smem) {
  TileModel md;
  tile_model_init(t, md, d, q, ql_param, smem);
  const int D = md.D, DD = D * D;
  const int ldp = 2 * D + 1, ldx = D + d + 1, lda = D + 1, ldz = D + d + 1, ldg = D + 1;
  double* PW = smem + TILE_MODEL_DOUBLES;    // D x 2D: [QL | F Uf] -> T
  double* X = PW + D * ldp;                  // (d+D) x (D+d); Uf = X[d:, d:d+D]
  double* A = X + (D + d) * ldx;             // D x D
  double* ZG = A + D * lda;                  // D x (D+d): [Z | G^T]
  double* G = ZG + D * ldz;                  // d x (D+1): G and, in column D, z
  double* Hs = G + d * ldg;                  // d x D
  TileVecs v;
  v.init(Hs + d * D, D, d);
  if (!lin.reg_sweeps) v.pb = nullptr;
  double* b = v.v0;
  double* eta = v.v1;
  const bool noisy = lin.R != nullptr;

  t.each(D * D, [&](int idx) {
    const int r = idx / D, c = idx - r * D;
    A[r * lda + c] = (r == c) ? 1.0 : 0.0;
    X[(long)(d + r) * ldx + d + c] = 0.0;
    ZG[r * ldz + c] = 0.0;
    if (c == 0) {
      b[r] = 0.0;
      eta[r] = 0.0;
    }
  });
  for (long k = k0; k < k1; ++k) {
   for (int ph = 0; ph < 3; ++ph) {
    TileSweep sw;
    sw.none();
    if (ph == 0) {
    tile_stage_lin(t, md, lin, k, Hs, v.cs, v.Rs);
    tile_set_step_model(md, lin, k);
    // predict: A <- F A, b <- F b, [QL | F Uf]
    if (md.Fd) {  // dense F: out of place through PW (free until the next each)
      t.each(D * (D + 1), [&](int idx) {
        const int r = idx / (D + 1), c = idx - r * (D + 1);
        if (c < D)
          PW[r * ldp + c] = tile_F_row(md, r, [&](int j) { return A[j * lda + c]; });
        else
          v.v2[r] = tile_F_row(md, r, [&](int j) { return b[j]; });
      });
      t.each(D * (D + 1), [&](int idx) {
        const int r = idx / (D + 1), c = idx - r * (D + 1);
        if (c < D)
          A[r * lda + c] = PW[r * ldp + c];
        else
          b[r] = v.v2[r];
      });
    } else {  // IWP: one thread per column, in place, rows ascending (row r only reads rows >= r of its block)
      t.each(D + 1, [&](int c) {
        for (int r = 0; r < D; ++r) {
          if (c < D)
            A[r * lda + c] = tile_F_row(md, r, [&](int j) { return A[j * lda + c]; });
          else
            b[r] = tile_F_row(md, r, [&](int j) { return b[j]; });
        }
      });
    }
    t.each(D * D, [&](int idx) {
      const int r = idx / D, c = idx - r * D;
      PW[r * ldp + c] = tile_QL(md, r, c);
      PW[r * ldp + D + c] = tile_F_row(md, r, [&](int j) { return X[(long)(d + j) * ldx + d + c]; });
    });
    sw.set(PW, D, 2 * D, ldp, D, D);
    } else if (ph == 1) {
    if (aggm && k == k1 - 1) {
      t.each(D * D, [&](int idx) {
        const int r = idx / D, c = idx - r * D;
        aggm[idx] = A[r * lda + c];
        aggm[DD + D + idx] = (c <= r) ? PW[r * ldp + c] : 0.0;
        aggm[2 * DD + 2 * D + idx] = (c <= r) ? ZG[r * ldz + c] : 0.0;
        if (c == 0) {
          aggm[DD + r] = b[r];
          aggm[2 * DD + D + r] = eta[r];
        }
      });
    }
    tile_update_build(t, md, PW, ldp, Hs, v.Rs, X);
    sw.set(X, d + D, noisy ? D + d : D, ldx, d, -1);
    } else {
    // G = SL^{-1} (H A) (d x D) and z = SL^{-1} (H b + c) in column D: products, then one thread per column solves
    t.each(d * (D + 1), [&](int idx) {
      const int a = idx / (D + 1), j = idx - a * (D + 1);
      G[a * ldg + j] = (j < D) ? tile_dot(0, D, 0.0, [&](int i) { return Hs[a * D + i]; },
                                          [&](int i) { return A[i * lda + j]; })
                               : tile_dot(0, D, v.cs[a], [&](int i) { return Hs[a * D + i]; },
                                          [&](int i) { return b[i]; });
    });
    t.each(D + 1, [&](int j) {
      for (int a = 0; a < d; ++a) {
        const double s = tile_dot(0, a, G[a * ldg + j], [&](int e) { return -X[(long)a * ldx + e]; },
                                  [&](int e) { return G[e * ldg + j]; });
        G[a * ldg + j] = s / X[(long)a * ldx + a];
      }
    });
    // A <- A - Kbar G ; b <- b - Kbar z ; eta <- eta - G^T z ; [Z | G^T]
    t.each(D * D, [&](int idx) {
      const int i = idx / D, j = idx - i * D;
      A[i * lda + j] = tile_dot(0, d, A[i * lda + j], [&](int a) { return -X[(long)(d + i) * ldx + a]; },
                                [&](int a) { return G[a * ldg + j]; });
      if (j < d) ZG[i * ldz + D + j] = G[j * ldg + i];
      if (j == 0) {
        b[i] = tile_dot(0, d, b[i], [&](int a) { return -X[(long)(d + i) * ldx + a]; },
                        [&](int a) { return G[a * ldg + D]; });
        eta[i] = tile_dot(0, d, eta[i], [&](int a) { return -G[a * ldg + i]; }, [&](int a) { return G[a * ldg + D]; });
      }
    });
    sw.set(ZG, D, D + d, ldz, D, D);
    }
    if (sw.M) tile_tria(t, sw.M, sw.R, sw.C, sw.ld, sw.npiv, sw.c0, v.diag, v.pb);
   }
  }
  t.each(D * D, [&](int idx) {
    const int r = idx / D, c = idx - r * D;
    agg[idx] = A[r * lda + c];
    agg[DD + D + idx] = X[(long)(d + r) * ldx + d + c];
    agg[2 * DD + 2 * D + idx] = (c <= r) ? ZG[r * ldz + c] : 0.0;
    if (c == 0) {
      agg[DD + r] = b[r];
      agg[2 * DD + D + r] = eta[r];
    }
  });
}

// ---------------------------------------------------------------------------------------------------------------
// filter phase 3
// ---------------------------------------------------------------------------------------------------------------
// state_in: packed (m, L) at the chunk start.  kern: per-step backward kernels, time-major, NE = D + 2 D^2 doubles per
// step: [g | E | Dk (lower)].  state_end: packed filtered state at the chunk end (L lower).  part: [nll, ssq_ref,
// ssq_proper].  fmeans / fchols (optional): filtered states k+1 in API layout.
POF_TDEV void tile_scan(const Team& t, int d, int q, const double* ql_param, const TileLin& lin, long k0, long k1,
                        const double* __restrict__ state_in, double* __restrict__ kern,
                        double* __restrict__ state_end, double* __restrict__ part, double* __restrict__ fmeans,
                        double* __restrict__ fchols, double* smem, const TileEks* eks = nullptr) {
  TileModel md;
  tile_model_init(t, md, d, q, ql_param, smem);
  const int D = md.D, DD = D * D, NE = D + 2 * DD;
  const int ldp = 2 * D + 1, ldx = D + d + 1;
  double* PW = smem + TILE_MODEL_DOUBLES;  // 2D x 2D: [[QL, F Uf],[0, Uf]] -> [[T, *],[Phi21, Phi22~]]
  double* X = PW + 2 * D * ldp;
  double* Hs = X + (D + d) * ldx;
  TileVecs v;
  v.init(Hs + d * D, D, d);
  if (!lin.reg_sweeps) v.pb = nullptr;
  double* m = v.v0;
  double* mp = v.v1;
  double* g = v.v2;
  const bool noisy = lin.R != nullptr;
  const double LOG_2PI = 1.8378770664093454835606594728112;

  t.each(D * D, [&](int idx) {
    const int r = idx / D, c = idx - r * D;
    X[(long)(d + r) * ldx + d + c] = state_in[D + idx];
    if (c == 0) m[r] = state_in[r];
    if (idx < 3) v.acc[idx] = 0.0;
  });
  // k == k1 is the epilogue: the end state goes to the smoother tree as (m, L) with L lower triangular
  for (long k = k0; k <= k1; ++k) {
   const int nph = (k < k1) ? 4 : 1;
   for (int ph = 0; ph < nph; ++ph) {
    TileSweep sw;
    sw.none();
    if (k == k1) {
      sw.set(X + (long)d * ldx + d, D, D, ldx, D, -1);
    } else if (ph == 0) {
    if (!eks) tile_stage_lin(t, md, lin, k, Hs, v.cs, v.Rs);
    tile_set_step_model(md, lin, k);
    t.each(D * D, [&](int idx) {
      const int r = idx / D, c = idx - r * D;
      PW[r * ldp + c] = tile_QL(md, r, c);
      PW[r * ldp + D + c] = tile_F_row(md, r, [&](int j) { return X[(long)(d + j) * ldx + d + c]; });
      PW[(D + r) * ldp + c] = 0.0;
      PW[(D + r) * ldp + D + c] = X[(long)(d + r) * ldx + d + c];
      if (c == 0) mp[r] = tile_F_row(md, r, [&](int j) { return m[j]; });
    });
    if (eks) tile_stage_lin_eks(t, md, *eks, lin.s0, lin.s1, mp, Hs, v.cs, v.Rs);
    sw.set(PW, 2 * D, 2 * D, ldp, D, D);
    } else if (ph == 1) {
    // E = Phi21 T^{-1} (row-wise back substitution, in place), then g = m - E (F m)
    t.each(D, [&](int r) {
      double* e = PW + (long)(D + r) * ldp;
      double gr = m[r];
      for (int j = D - 1; j >= 0; --j) {
        double s = e[j];
        s = tile_dot(j + 1, D, s, [&](int i) { return -e[i]; }, [&](int i) { return PW[(long)i * ldp + j]; });
        s /= PW[(long)j * ldp + j];
        e[j] = s;
        gr = fma(-s, mp[j], gr);
      }
      g[r] = gr;
    });
    // Dk = tria(Phi22~)
    sw.set(PW + (long)D * ldp + D, D, D, ldp, D, -1);
    } else if (ph == 2) {
    {
      double* kp = kern + k * NE;
      t.each(NE, [&](int idx) {
        double val;
        if (idx < D) {
          val = g[idx];
        } else if (idx < D + DD) {
          const int r = (idx - D) / D, c = (idx - D) - r * D;
          val = PW[(long)(D + r) * ldp + c];
        } else {
          const int r = (idx - D - DD) / D, c = (idx - D - DD) - r * D;
          val = (c <= r) ? PW[(long)(D + r) * ldp + D + c] : 0.0;
        }
        kp[idx] = val;
      });
    }
    tile_update_build(t, md, PW, ldp, Hs, v.Rs, X);
    sw.set(X, d + D, noisy ? D + d : D, ldx, d, -1);
    } else {
    t.each(d, [&](int a) {
      double s = v.cs[a];
      s = tile_dot(0, D, s, [&](int i) { return Hs[a * D + i]; }, [&](int i) { return mp[i]; });
      v.y[a] = s;
    });
    // innovation statistics: nll = -log N(y; 0, SL SL^T) (pof/utils.py:22-30), ssq_ref = |SL^{-T} y|^2 (utils.py:110-112
    // solves with the transpose), ssq_proper = |SL^{-1} y|^2
    t.each(1, [&](int) {
      double zz = 0.0, lg = 0.0, ww = 0.0;
      for (int a = 0; a < d; ++a) {
        double s = v.y[a];
        s = tile_dot(0, a, s, [&](int e) { return -X[(long)a * ldx + e]; }, [&](int e) { return v.z[e]; });
        s /= X[(long)a * ldx + a];
        v.z[a] = s;
        zz = fma(s, s, zz);
        lg += log(fabs(X[(long)a * ldx + a]));
      }
      for (int a = d - 1; a >= 0; --a) {
        double s = v.y[a];
        s = tile_dot(a + 1, d, s, [&](int e) { return -X[(long)e * ldx + a]; }, [&](int e) { return v.w[e]; });
        s /= X[(long)a * ldx + a];
        v.w[a] = s;
        ww = fma(s, s, ww);
      }
      v.acc[0] += 0.5 * zz + lg + 0.5 * d * LOG_2PI;
      v.acc[1] += ww;
      v.acc[2] += zz;
    });
    t.each(D, [&](int i) {
      double s = mp[i];
      s = tile_dot(0, d, s, [&](int a) { return -X[(long)(d + i) * ldx + a]; }, [&](int a) { return v.z[a]; });
      m[i] = s;
    });
    if (fmeans) {
      t.each(D * D, [&](int idx) {
        const int r = idx / D, c = idx - r * D;
        fchols[(k + 1) * DD + idx] = X[(long)(d + r) * ldx + d + c];
        if (c == 0) fmeans[(k + 1) * D + r] = m[r];
      });
    }
    }
    if (sw.M) tile_tria(t, sw.M, sw.R, sw.C, sw.ld, sw.npiv, sw.c0, v.diag, v.pb);
   }
  }
  t.each(D * D, [&](int idx) {
    const int r = idx / D, c = idx - r * D;
    state_end[D + idx] = (c <= r) ? X[(long)(d + r) * ldx + d + c] : 0.0;
    if (c == 0) state_end[r] = m[r];
    if (idx < 3) part[idx] = v.acc[idx];
  });
}

// ---------------------------------------------------------------------------------------------------------------
// smoother phase 3
// ---------------------------------------------------------------------------------------------------------------
// seed: packed smoothed state (m, L lower) at time k1.  Writes the smoothed states t in [k0, k1) (and t = n when
// `last`) in API layout, factors scaled by cscale (pof/step.py:42-44).  part: [obj, #mean entries not close].
POF_TDEV void tile_smooth(const Team& t, int d, int q, const double* ql_param, const TileLin& lin, long k0, long k1,
                          bool last,
                          bool emit_t0, const double* __restrict__ seed, const double* __restrict__ kern,
                          double cscale, double* __restrict__ means, double* __restrict__ chols,
                          double* __restrict__ part, double* smem) {
  TileModel md;
  tile_model_init(t, md, d, q, ql_param, smem);
  const int D = md.D, DD = D * D, NE = D + 2 * DD, Q1 = md.Q1;
  const int lds = 2 * D + 1, lde = D + 1;
  double* SW = smem + TILE_MODEL_DOUBLES;  // D x 2D: [Dk | E L] -> new L
  double* Es = SW + D * lds;               // D x D
  double* Ls = Es + D * lde;               // D x D (lower)
  TileVecs v;
  v.init(Ls + D * lde, D, d);
  if (!lin.reg_sweeps) v.pb = nullptr;
  double* m = v.v0;
  double* mn = v.v1;
  double* fm = v.v2;
  double* bad = v.v3;  // per-row counters of the isclose test
  double* rs = v.diag + 2 * D;  // QL^{-1} residual of the objective

  t.each(D * D, [&](int idx) {
    const int r = idx / D, c = idx - r * D;
    Ls[r * lde + c] = (c <= r) ? seed[D + idx] : 0.0;
    if (c == 0) {
      m[r] = seed[r];
      bad[r] = 0.0;
    }
    if (idx == 0) v.acc[0] = 0.0;
  });
  // write one smoothed state; counts the mean entries failing isclose(old, new, rtol=1e-13, atol=1e-8)
  // (pof/convergence_criteria.py:9; NaN is never close)
  auto emit = [&](long ti) {
    t.each(D * D, [&](int idx) {
      const int r = idx / D, c = idx - r * D;
      if (chols) chols[ti * DD + idx] = (c <= r) ? cscale * Ls[r * lde + c] : 0.0;
      if (c == 0) {
        const double old = means[ti * D + r];
        const double nw = m[r];
        const bool close = fabs(old - nw) <= (1e-8 + 1e-13 * fabs(nw));
        bad[r] += close ? 0.0 : 1.0;
        means[ti * D + r] = nw;
      }
    });
  };
  if (last) emit(k1);
  for (long k = k1 - 1; k >= k0; --k) {
   for (int ph = 0; ph < 2; ++ph) {
    TileSweep sw;
    sw.none();
    const double* kp = kern + k * NE;
    tile_set_step_model(md, lin, k);
    if (ph == 0) {
    t.each(D * D, [&](int idx) {
      const int r = idx / D, c = idx - r * D;
      Es[r * lde + c] = kp[D + idx];
      SW[r * lds + c] = kp[D + DD + idx];
      if (c == 0) fm[r] = tile_F_row(md, r, [&](int j) { return m[j]; });
    });
    // [Dk | E L], mn = g + E m
    t.each(D * D + D, [&](int idx) {
      if (idx < DD) {
        const int a = idx / D, j = idx - a * D;
        double s = 0.0;
        s = tile_dot(j, D, s, [&](int i) { return Es[a * lde + i]; }, [&](int i) { return Ls[i * lde + j]; });
        SW[a * lds + D + j] = s;
      } else {
        const int a = idx - DD;
        double s = kp[a];
        s = tile_dot(0, D, s, [&](int i) { return Es[a * lde + i]; }, [&](int i) { return m[i]; });
        mn[a] = s;
      }
    });
    sw.set(SW, D, 2 * D, lds, D, D);
    } else {
    // objective increment |QL^{-1}(m_k - F m_{k+1})|^2 (reference's swapped-argument form, smoother.py:20): one thread
    // per block of the block-diagonal QL; new state
    t.each(d, [&](int b) {
      double o = 0.0;
      if (md.Qd) {  // dense QL: one forward substitution over the whole state (iteration 0), nothing for the others
        if (b == 0) {
          for (int i = 0; i < D; ++i) {
            double s = mn[i] - fm[i];
            s = tile_dot(0, i, s, [&](int j) { return -md.Qd[i * D + j]; }, [&](int j) { return rs[j]; });
            s /= md.Qd[i * D + i];
            rs[i] = s;
            o = fma(s, s, o);
          }
        }
      } else {
        for (int i = 0; i < Q1; ++i) {
          double s = mn[b * Q1 + i] - fm[b * Q1 + i];
          s = tile_dot(0, i, s, [&](int j) { return -md.ql[i * Q1 + j]; }, [&](int j) { return rs[b * Q1 + j]; });
          s /= md.ql[i * Q1 + i];
          rs[b * Q1 + i] = s;
          o = fma(s, s, o);
        }
      }
      v.w[b] = o;
    });
    t.each(D * D, [&](int idx) {
      const int r = idx / D, c = idx - r * D;
      Ls[r * lde + c] = (c <= r) ? SW[r * lds + c] : 0.0;
      if (c == 0) m[r] = mn[r];
      if (idx == 0) {
        double o = v.acc[0];
        for (int b = 0; b < d; ++b) o += v.w[b];
        v.acc[0] = o;
      }
    });
    if (k > 0 || emit_t0) emit(k);
    }
    if (sw.M) tile_tria(t, sw.M, sw.R, sw.C, sw.ld, sw.npiv, sw.c0, v.diag, v.pb);
   }
  }
  t.each(1, [&](int) {
    double nb = 0.0;
    for (int r = 0; r < D; ++r) nb += bad[r];
    part[0] = v.acc[0];
    part[1] = nb;
  });
}

// Sequential extended Kalman smoother on ONE CTA (reference pof/sequential_filtsmooth/__init__.py:5-10): the filter
// relinearised at every predicted mean over the whole grid, then the RTS recursion.  x0: packed (m, L).  kern: (n, NE)
// scratch, state_end: D + D^2 scratch.  sums: [sum -loglik, ssq_ref sum, ssq_proper sum, obj, (unused count)].
POF_TDEV void tile_seq_eks(const Team& t, int d, int q, const double* ql_param, const TileLin& lin, const TileEks& eks,
                           long n, const double* __restrict__ x0, double* __restrict__ kern,
                           double* __restrict__ state_end, double* __restrict__ means, double* __restrict__ chols,
                           double* __restrict__ sums, double* smem) {
  tile_scan(t, d, q, ql_param, lin, 0, n, x0, kern, state_end, sums, nullptr, nullptr, smem, &eks);
  tile_smooth(t, d, q, ql_param, lin, 0, n, true, true, state_end, kern, 1.0, means, chols, sums + 3, smem);
}

// ---------------------------------------------------------------------------------------------------------------
// tree operators, one CTA per node; operands are read from global memory (L2), the work arrays Xi (2D x 2D) and
// W (D x 2D) live in shared memory.  Algebra as in pof_coop.cuh (equivalent to filter.py:117-142):
//   Y = U1 Xi11^{-T}    G = I - Y Xi21^T    A = A2 G A1    b = A2 G (b1 + U1 U1^T eta2) + b2    U = tria([A2 Y, U2])
//   eta = A1^T G^T (eta2 - Z2 Z2^T b1) + eta1                                                   Z = tria([A1^T Xi22, Z1])
// ---------------------------------------------------------------------------------------------------------------
struct TileTreeWs {
  double *Xi, *W, *t0, *t1, *t2, *t3, *diag, *pb;
  int ldx;
  POF_TDEV TileTreeWs(double* smem, int D) {
    ldx = 2 * D + 1;
    Xi = smem;
    W = Xi + 2 * D * ldx;
    t0 = W + D * ldx;
    t1 = t0 + D;
    t2 = t1 + D;
    t3 = t2 + D;
    diag = t3 + D;
    pb = nullptr;  // the tree operators always run the shared-memory sweep (2D columns; a few hundred nodes per pass)
  }
};

// e1: earlier element (packed filtering element, or packed state (m, L) if state_mode), e2: later element.
// out: packed filtering element, or packed state if state_mode.  out must not alias e1 / e2.
POF_TDEV void tile_filter_combine(const Team& t, int D, const double* e1, const double* e2,
                                  double* out, double* smem, bool state_mode) {
  TileTreeWs s(smem, D);
  const int DD = D * D, ldx = s.ldx;
  const double* A1 = state_mode ? nullptr : e1;
  const double* b1 = state_mode ? e1 : e1 + DD;
  const double* U1 = state_mode ? e1 + D : e1 + DD + D;
  const double* n1 = state_mode ? nullptr : e1 + 2 * DD + D;
  const double* Z1 = state_mode ? nullptr : e1 + 2 * DD + 2 * D;
  const double* A2 = e2;
  const double* b2 = e2 + DD;
  const double* U2 = e2 + DD + D;
  const double* n2 = e2 + 2 * DD + D;
  const double* Z2 = e2 + 2 * DD + 2 * D;
  double* Xi = s.Xi;
  double* W = s.W;
  // Xi = [[U1^T Z2, I],[Z2, 0]]
  t.each(DD, [&](int idx) {
    const int r = idx / D, c = idx - r * D;
    double v = 0.0;
    v = tile_dot(0, D, v, [&](int k) { return U1[k * D + r]; }, [&](int k) { return Z2[k * D + c]; });
    Xi[r * ldx + c] = v;
    Xi[r * ldx + D + c] = (r == c) ? 1.0 : 0.0;
    Xi[(D + r) * ldx + c] = Z2[idx];
    Xi[(D + r) * ldx + D + c] = 0.0;
  });
  tile_tria(t, Xi, 2 * D, 2 * D, ldx, state_mode ? D : 2 * D, -1, s.diag, s.pb);
  // Y = U1 Xi11^{-T} (row r solves y Xi11^T = u_r), kept in the dead upper-right block of Xi
  double* Y = Xi + D;  // Y(r, j) = Y[r * ldx + j]
  t.each(D, [&](int r) {
    for (int j = 0; j < D; ++j) {
      double acc = U1[r * D + j];
      acc = tile_dot(0, j, acc, [&](int i) { return -Y[r * ldx + i]; }, [&](int i) { return Xi[j * ldx + i]; });
      Y[r * ldx + j] = acc / Xi[j * ldx + j];
    }
  });
  // G = I - Y Xi21^T in W[:, 0:D];  t1 = U1^T eta2;  t3 = Z2^T b1
  double* G = W;
  t.each(DD + 2 * D, [&](int idx) {
    if (idx < DD) {
      const int r = idx / D, c = idx - r * D;
      double v = 0.0;
      v = tile_dot(0, D, v, [&](int k) { return Y[r * ldx + k]; }, [&](int k) { return Xi[(D + c) * ldx + k]; });
      G[r * ldx + c] = ((r == c) ? 1.0 : 0.0) - v;
    } else if (idx < DD + D) {
      const int i = idx - DD;
      double acc = 0.0;
      acc = tile_dot(0, D, acc, [&](int k) { return U1[k * D + i]; }, [&](int k) { return n2[k]; });
      s.t1[i] = acc;
    } else {
      const int i = idx - DD - D;
      double acc = 0.0;
      acc = tile_dot(0, D, acc, [&](int k) { return Z2[k * D + i]; }, [&](int k) { return b1[k]; });
      s.t3[i] = acc;
    }
  });
  // t0 = b1 + U1 t1 ;  t2 = eta2 - Z2 t3
  t.each(2 * D, [&](int idx) {
    if (idx < D) {
      double acc = b1[idx];
      acc = tile_dot(0, D, acc, [&](int k) { return U1[idx * D + k]; }, [&](int k) { return s.t1[k]; });
      s.t0[idx] = acc;
    } else {
      const int i = idx - D;
      double acc = n2[i];
      acc = tile_dot(0, D, acc, [&](int k) { return -Z2[i * D + k]; }, [&](int k) { return s.t3[k]; });
      s.t2[i] = acc;
    }
  });
  // t1 = G t0 ; t3 = G^T t2 ; P = G A1 in W[:, D:2D]
  double* P = W + D;
  t.each((state_mode ? 0 : DD) + 2 * D, [&](int idx0) {
    if (idx0 < 2 * D) {
      if (idx0 < D) {
        double acc = 0.0;
        acc = tile_dot(0, D, acc, [&](int k) { return G[idx0 * ldx + k]; }, [&](int k) { return s.t0[k]; });
        s.t1[idx0] = acc;
      } else {
        const int i = idx0 - D;
        double acc = 0.0;
        acc = tile_dot(0, D, acc, [&](int k) { return G[k * ldx + i]; }, [&](int k) { return s.t2[k]; });
        s.t3[i] = acc;
      }
    } else {
      const int idx = idx0 - 2 * D;
      const int r = idx / D, c = idx - r * D;
      double v = 0.0;
      v = tile_dot(0, D, v, [&](int k) { return G[r * ldx + k]; }, [&](int k) { return A1[k * D + c]; });
      P[r * ldx + c] = v;
    }
  });
  // b = A2 t1 + b2 ;  eta = A1^T t3 + eta1 ;  A = A2 P
  double* ob = state_mode ? out : out + DD;
  t.each((state_mode ? 0 : DD + D) + D, [&](int idx0) {
    if (idx0 < D) {
      double acc = b2[idx0];
      acc = tile_dot(0, D, acc, [&](int k) { return A2[idx0 * D + k]; }, [&](int k) { return s.t1[k]; });
      ob[idx0] = acc;
    } else if (idx0 < 2 * D) {
      const int i = idx0 - D;
      double acc = n1[i];
      acc = tile_dot(0, D, acc, [&](int k) { return A1[k * D + i]; }, [&](int k) { return s.t3[k]; });
      out[2 * DD + D + i] = acc;
    } else {
      const int idx = idx0 - 2 * D;
      const int r = idx / D, c = idx - r * D;
      double v = 0.0;
      v = tile_dot(0, D, v, [&](int k) { return A2[r * D + k]; }, [&](int k) { return P[k * ldx + c]; });
      out[idx] = v;
    }
  });
  // U = tria([A2 Y, U2])
  t.each(DD, [&](int idx) {
    const int r = idx / D, c = idx - r * D;
    double v = 0.0;
    v = tile_dot(0, D, v, [&](int k) { return A2[r * D + k]; }, [&](int k) { return Y[k * ldx + c]; });
    W[r * ldx + c] = v;
    W[r * ldx + D + c] = U2[idx];
  });
  tile_tria(t, W, D, 2 * D, ldx, D, -1, s.diag, s.pb);
  double* oU = state_mode ? out + D : out + DD + D;
  t.each(DD, [&](int idx) {
    const int r = idx / D, c = idx - r * D;
    oU[idx] = (c <= r) ? W[r * ldx + c] : 0.0;
  });
  if (state_mode) return;
  // Z = tria([A1^T Xi22, Z1])   (Xi22 lower triangular)
  t.each(DD, [&](int idx) {
    const int r = idx / D, c = idx - r * D;
    double v = 0.0;
    v = tile_dot(c, D, v, [&](int k) { return A1[k * D + r]; }, [&](int k) { return Xi[(D + k) * ldx + D + c]; });
    W[r * ldx + c] = v;
    W[r * ldx + D + c] = Z1[idx];
  });
  tile_tria(t, W, D, 2 * D, ldx, D, -1, s.diag, s.pb);
  double* oZ = out + 2 * DD + 2 * D;
  t.each(DD, [&](int idx) {
    const int r = idx / D, c = idx - r * D;
    oZ[idx] = (c <= r) ? W[r * ldx + c] : 0.0;
  });
}

// e1: LATER element (packed smoothing element [g | E | Dm], or packed state if state_mode), e2: EARLIER element.
// g = E2 g1 + g2 ; E = E2 E1 ; Dm = tria([E2 D1, D2])   (smoother.py:53-63)
POF_TDEV void tile_smooth_combine(const Team& t, int D, const double* e1, const double* e2,
                                  double* out, double* smem, bool state_mode) {
  TileTreeWs s(smem, D);
  const int DD = D * D, ldx = s.ldx;
  const double* g1 = e1;
  const double* E1 = state_mode ? nullptr : e1 + D;
  const double* D1 = state_mode ? e1 + D : e1 + D + DD;
  const double* g2 = e2;
  const double* E2 = e2 + D;
  const double* D2 = e2 + D + DD;
  double* W = s.W;
  t.each(DD + D + (state_mode ? 0 : DD), [&](int idx0) {
    if (idx0 < DD) {
      const int r = idx0 / D, c = idx0 - r * D;
      double v = 0.0;
      v = tile_dot(0, D, v, [&](int k) { return E2[r * D + k]; }, [&](int k) { return D1[k * D + c]; });
      W[r * ldx + c] = v;
      W[r * ldx + D + c] = D2[idx0];
    } else if (idx0 < DD + D) {
      const int i = idx0 - DD;
      double acc = g2[i];
      acc = tile_dot(0, D, acc, [&](int k) { return E2[i * D + k]; }, [&](int k) { return g1[k]; });
      out[i] = acc;
    } else {
      const int idx = idx0 - DD - D;
      const int r = idx / D, c = idx - r * D;
      double v = 0.0;
      v = tile_dot(0, D, v, [&](int k) { return E2[r * D + k]; }, [&](int k) { return E1[k * D + c]; });
      out[D + idx] = v;
    }
  });
  tile_tria(t, W, D, 2 * D, ldx, D, -1, s.diag, s.pb);
  double* oD = state_mode ? out + D : out + D + DD;
  t.each(DD, [&](int idx) {
    const int r = idx / D, c = idx - r * D;
    oD[idx] = (c <= r) ? W[r * ldx + c] : 0.0;
  });
}

// The smoothing element (g, E, Dm) of a whole chunk from the filtered state st = (m, L) at the chunk start and the
// chunk's filtering element e2 = (A, b, U, eta, Z) taken before its last measurement update (see pof_treelane.cuh):
//   Xi = tria([[L^T Z, I],[Z, 0]]),  Y = L Xi11^{-T},  G = I - Y Xi21^T,  m' = G (m + L L^T eta)
//   tria([[A Y, U],[Y, 0]]) = [[Phi11, 0],[Phi21, Phi22]],  E = Phi21 Phi11^{-1},  g = m' - E (A m' + b),  Dm = Phi22
POF_TDEV void tile_chunk_kernel(const Team& t, int D, const double* st, const double* e2,
                                double* out, double* smem) {
  TileTreeWs s(smem, D);
  const int DD = D * D, ldx = s.ldx;
  const double* m1 = st;
  const double* L1 = st + D;
  const double* A2 = e2;
  const double* b2 = e2 + DD;
  const double* U2 = e2 + DD + D;
  const double* n2 = e2 + 2 * DD + D;
  const double* Z2 = e2 + 2 * DD + 2 * D;
  double* Xi = s.Xi;
  double* W = s.W;
  t.each(DD + D, [&](int idx) {
    if (idx < DD) {
      const int r = idx / D, c = idx - r * D;
      double v = 0.0;
      v = tile_dot(0, D, v, [&](int k) { return L1[k * D + r]; }, [&](int k) { return Z2[k * D + c]; });
      Xi[r * ldx + c] = v;
      Xi[r * ldx + D + c] = (r == c) ? 1.0 : 0.0;
      Xi[(D + r) * ldx + c] = Z2[idx];
      Xi[(D + r) * ldx + D + c] = 0.0;
    } else {
      const int i = idx - DD;
      double acc = 0.0;
      acc = tile_dot(0, D, acc, [&](int k) { return L1[k * D + i]; }, [&](int k) { return n2[k]; });
      s.t1[i] = acc;  // L^T eta
    }
  });
  tile_tria(t, Xi, 2 * D, 2 * D, ldx, D, -1, s.diag, s.pb);
  // Y = L Xi11^{-T} in W[:, 0:D];  t0 = m + L t1
  double* Y = W;
  t.each(2 * D, [&](int idx) {
    if (idx < D) {
      const int r = idx;
      for (int j = 0; j < D; ++j) {
        double acc = L1[r * D + j];
        acc = tile_dot(0, j, acc, [&](int i) { return -Y[r * ldx + i]; }, [&](int i) { return Xi[j * ldx + i]; });
        Y[r * ldx + j] = acc / Xi[j * ldx + j];
      }
    } else {
      const int i = idx - D;
      double acc = m1[i];
      acc = tile_dot(0, D, acc, [&](int k) { return L1[i * D + k]; }, [&](int k) { return s.t1[k]; });
      s.t0[i] = acc;
    }
  });
  // G = I - Y Xi21^T in W[:, D:2D]
  double* G = W + D;
  t.each(DD, [&](int idx) {
    const int r = idx / D, c = idx - r * D;
    double v = 0.0;
    v = tile_dot(0, D, v, [&](int k) { return Y[r * ldx + k]; }, [&](int k) { return Xi[(D + c) * ldx + k]; });
    G[r * ldx + c] = ((r == c) ? 1.0 : 0.0) - v;
  });
  // m' = G t0 -> t2
  t.each(D, [&](int i) {
    double acc = 0.0;
    acc = tile_dot(0, D, acc, [&](int k) { return G[i * ldx + k]; }, [&](int k) { return s.t0[k]; });
    s.t2[i] = acc;
  });
  // v = A m' + b -> t3 ;  Phi = [[A Y, U],[Y, 0]] overwrites Xi (dead now)
  t.each(DD + D, [&](int idx) {
    if (idx < DD) {
      const int r = idx / D, c = idx - r * D;
      double v = 0.0;
      v = tile_dot(0, D, v, [&](int k) { return A2[r * D + k]; }, [&](int k) { return Y[k * ldx + c]; });
      Xi[r * ldx + c] = v;
      Xi[r * ldx + D + c] = U2[idx];
      Xi[(D + r) * ldx + c] = Y[r * ldx + c];
      Xi[(D + r) * ldx + D + c] = 0.0;
    } else {
      const int i = idx - DD;
      double acc = b2[i];
      acc = tile_dot(0, D, acc, [&](int k) { return A2[i * D + k]; }, [&](int k) { return s.t2[k]; });
      s.t3[i] = acc;
    }
  });
  tile_tria(t, Xi, 2 * D, 2 * D, ldx, 2 * D, -1, s.diag, s.pb);
  // E = Phi21 Phi11^{-1} (row-wise back substitution, in place); g = m' - E v; outputs
  t.each(D, [&](int r) {
    double* e = Xi + (long)(D + r) * ldx;
    double gr = s.t2[r];
    for (int j = D - 1; j >= 0; --j) {
      double acc = e[j];
      acc = tile_dot(j + 1, D, acc, [&](int i) { return -e[i]; }, [&](int i) { return Xi[(long)i * ldx + j]; });
      acc /= Xi[(long)j * ldx + j];
      e[j] = acc;
      gr = fma(-acc, s.t3[j], gr);
    }
    out[r] = gr;
  });
  t.each(DD, [&](int idx) {
    const int r = idx / D, c = idx - r * D;
    out[D + idx] = Xi[(long)(D + r) * ldx + c];
    out[D + DD + idx] = (c <= r) ? Xi[(long)(D + r) * ldx + D + c] : 0.0;
  });
}

}  // namespace pof
