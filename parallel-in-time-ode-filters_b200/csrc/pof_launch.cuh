// Launch interface between the (d, q)-templated leaf kernels (compiled in pof_leaf_d*.cu, one translation unit
// per ODE dimension so that nvcc builds them in parallel) and the C ABI in pof_api.cu.
#pragma once
#include <cuda_runtime.h>

#include <mutex>

#include "pof_real.cuh"

namespace POF_NS {

// Raise a kernel's dynamic shared-memory limit ONCE per (kernel, device): cudaFuncSetAttribute is a driver call of
// several microseconds, and calling it on every launch made the tiny tree kernels launch-bound.
struct SmemCache {
  static constexpr int CAP = 1024;
  const void* fn[CAP];
  int dev[CAP], bytes[CAP], n = 0;
};
inline SmemCache& smem_cache() {
  static SmemCache c;
  return c;
}
inline std::mutex& smem_cache_mutex() {
  static std::mutex m;
  return m;
}
// max_carveout: prefer the largest shared-memory carve-out for this kernel.  Kernels of different sweeps co-reside
// on an SM (the smoother's up-sweep runs next to the filter scan) only if the carve-out chosen for the resident one
// leaves room for the other; kernels that run alone keep the default (a larger L1 is worth ~3 % to them).
template <class K>
inline cudaError_t ensure_smem(K kernel, int bytes, bool max_carveout = false) {
  if (bytes <= 48 * 1024 && !max_carveout) return cudaSuccess;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(smem_cache_mutex());
  SmemCache& c = smem_cache();
  const void* key = (const void*)kernel;
  for (int i = 0; i < c.n; ++i)
    if (c.fn[i] == key && c.dev[i] == dev && c.bytes[i] >= bytes) return cudaSuccess;
  if (max_carveout)
    (void)cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaError_t e = bytes > 48 * 1024
                      ? cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)
                      : cudaSuccess;
  if (e == cudaSuccess && c.n < SmemCache::CAP) {
    c.fn[c.n] = key;
    c.dev[c.n] = dev;
    c.bytes[c.n] = bytes;
    ++c.n;
  }
  return e;
}

struct QLParam {
  real v[36];  // (q+1) x (q+1) row-major, q <= 5; lives in the kernel-parameter constant bank
};

struct LeafArgs {
  long n;    // local number of steps
  long L;    // chunk length
  long CS;   // number of chunks
  const real* H;  // dense linearisation (n,d,D), (n,d) ...
  const real* c;
  const real* Jc;  // ... or compact: per step [J_f (d x d) | c (d)], H = E1 - J_f E0 rebuilt on load (Jc != null)
  real s0, s1;     // Nordsieck scalings: E0 = s0 e_0^T, E1 = s1 e_1^T per block
  QLParam ql;
  int d, q;          // runtime dimensions (the tile family is not templated on them)
  const real* R;   // observation-noise factors cholR (n,d,d), or null = noiseless (tile family only)
  const real* F;   // general per-step transition model (n,D,D) x 2 (QLd lower triangular), or null = the
  const real* QLd; // preconditioned IWP described by ql (tile family only)
  int tile_reg;      // tile family: register-resident Householder sweeps (default; 0 with POF_F_TILE_SMEM_QR)
};

struct LeafLaunch {
  // faggm (may be null): the chunk's filtering element BEFORE its last measurement update (lane kernels only)
  cudaError_t (*fold)(cudaStream_t, const LeafArgs&, real* fagg, real* faggm);
  // sagg null: do not compose the chunk's smoothing element inside the scan (lane kernels only)
  cudaError_t (*scan)(cudaStream_t, const LeafArgs&, const real* fin, real* kern, real* sagg, real* send,
                      real* part, real* fmeans, real* fchols);
  cudaError_t (*smooth)(cudaStream_t, const LeafArgs&, const real* sin, const real* kern, int emit_t0,
                        const real* cscale, real* means, real* chols, real* part2);
  // sequential extended Kalman smoother relinearised at the predicted mean (one thread; baseline path), or null
  cudaError_t (*seq_eks)(cudaStream_t, const LeafArgs&, int ivp_id, const real* params8, const real* x0,
                         real* kern, real* means, real* chols, real* part);
  int chunks_per_warp;  // 32 / G for the lane-cooperative kernels
  int has_pre_update;   // 1 if fold emits faggm (the chunk's element before its last update); both families do
  int is_tile;          // 1 for the CTA-per-chunk large-state family (pof_tile.cu): chunks_per_warp is 0 there
};

#ifndef POF_F32
// one-thread sequential EKS (pof_seq_kernels.cuh): only `seq_eks` is set; nullptr if q is not compiled in
const LeafLaunch* seq_launch_d1(int q);
const LeafLaunch* seq_launch_d2(int q);
const LeafLaunch* seq_launch_d3(int q);
const LeafLaunch* seq_launch_d4(int q);
// two rows per lane (pof_lane2.cuh); null where not instantiated
const LeafLaunch* lane2_launch_d1(int q);
const LeafLaunch* lane2_launch_d2(int q);
const LeafLaunch* lane2_launch_d3(int q);
const LeafLaunch* lane2_launch_d4(int q);

// large-state family (pof_tile.cu): CTA per chunk / per tree node, runtime (d, q); D limited by shared memory
bool tile_supported(int d, int q);
bool tile_tree_supported(int D);
int tile_ctas_per_sm(int d, int q);
const LeafLaunch* tile_leaf_launch();
cudaError_t tile_fup(cudaStream_t, int D, const real* child, long nchild, real* parent, long nparent);
cudaError_t tile_fdown(cudaStream_t, int D, const real* pin, long nparent, const real* cagg, long nchild,
                       real* cin);
cudaError_t tile_sup(cudaStream_t, int D, const real* child, long nchild, real* parent, long nparent);
cudaError_t tile_sdown(cudaStream_t, int D, const real* pin, long nparent, const real* cagg, long nchild,
                       real* cin);
cudaError_t tile_chunkk(cudaStream_t, int D, const real* fin, const real* faggm, real* sagg, long CS);
cudaError_t tile_fcomb(cudaStream_t, int D, long n, const real* e1, const real* e2, real* out);
cudaError_t tile_scomb(cudaStream_t, int D, long n, const real* e1, const real* e2, real* out);
cudaError_t tile_fchain(cudaStream_t, int D, int count, const real* state_in, const real* elems,
                        real* state_out, real* scratch);
cudaError_t tile_schain(cudaStream_t, int D, int count, const real* state_in, const real* elems,
                        real* state_out, real* scratch);

#else
// fp32 build: only the register-resident family exists (the large-state tile kernels and the one-thread sequential EKS
// are fp64 only)
const LeafLaunch* lane2_launch_d1(int q);
const LeafLaunch* lane2_launch_d2(int q);
const LeafLaunch* lane2_launch_d3(int q);
const LeafLaunch* lane2_launch_d4(int q);
inline bool tile_supported(int, int) { return false; }
inline bool tile_tree_supported(int) { return false; }
inline int tile_ctas_per_sm(int, int) { return 1; }
inline const LeafLaunch* tile_leaf_launch() { return nullptr; }
inline cudaError_t tile_fup(cudaStream_t, int, const real*, long, real*, long) { return cudaErrorNotSupported; }
inline cudaError_t tile_fdown(cudaStream_t, int, const real*, long, const real*, long, real*) { return cudaErrorNotSupported; }
inline cudaError_t tile_sup(cudaStream_t, int, const real*, long, real*, long) { return cudaErrorNotSupported; }
inline cudaError_t tile_sdown(cudaStream_t, int, const real*, long, const real*, long, real*) { return cudaErrorNotSupported; }
inline cudaError_t tile_chunkk(cudaStream_t, int, const real*, const real*, real*, long) { return cudaErrorNotSupported; }
inline cudaError_t tile_fcomb(cudaStream_t, int, long, const real*, const real*, real*) { return cudaErrorNotSupported; }
inline cudaError_t tile_scomb(cudaStream_t, int, long, const real*, const real*, real*) { return cudaErrorNotSupported; }
#endif

// One whole tree sweep (an up-sweep and/or a down-sweep over the chunk carries) as ONE kernel without level barriers:
// a DATAFLOW schedule.  Work items (one associative combine each) are numbered in level order; every warp draws the
// next ticket from an atomic counter, waits until the flags of the nodes its items depend on are set (acquire loads),
// computes, and publishes its own nodes' flags (release).  An item only depends on items with SMALLER numbers, and
// those tickets are held by warps that are already running, so the schedule cannot deadlock whatever the residency.
// Compared with one launch per level (~10 us per level: launch gap + cold instruction cache + one combine's latency)
// a level costs one warm combine plus an L2 round trip for the flag.
struct FlowArgs {
  static constexpr int MAXL = 40;
  enum Kind { UP = 0, ROOT = 1, DOWN = 2, DOWN_E = 3 };
  int nlev;                 // levels of the tree (level 0 = chunks)
  long off[MAXL], sz[MAXL];
  int nseg;                 // segments in execution order: segment j has seg_count[j] independent items
  int seg_kind[2 * MAXL + 2];
  int seg_level[2 * MAXL + 2];  // UP: the level that is built (from level - 1); DOWN: the level whose states are known
  long seg_count[2 * MAXL + 2];
  long seg_begin[2 * MAXL + 3];  // filled by the launcher: item offsets, every segment padded to whole tickets (a
                                 // ticket never straddles two segments: its items must not depend on each other)
  int up_lo, up_hi;         // levels built by UP segments of THIS launch (their flags are waited for); elements of
                            // any other level were complete before the launch
  real* agg;              // elements per node (filtering: 3D^2+2D doubles, smoothing: 2D^2+D)
  real* st;               // states per node (D + D^2 doubles): incoming filtered / outgoing smoothed states
  real* sx;               // smoother, element-form down-sweep (DOWN_E): per node the aggregate of everything LATER
  const real* root_m;     // state of the root node for the down-sweep (ROOT item): mean (D), factor (D x D);
  const real* root_L;     // null with DOWN_E segments: the root's "later" aggregate is the identity element
  unsigned* flag_up;        // per node: element complete   } zeroed by a stream-ordered memset before the launch
  unsigned* flag_dn;        // per node: state complete     }
  unsigned* ticket;         // the ticket counter           }
};

struct ExchangeArgs {
  int rank, world;
  const real* gathered;   // world payloads, `stride` doubles each
  long stride;
  const real* x0_mean;    // filter: initial state
  const real* x0_chol;
  real* state_out;        // D + D*D: incoming filtered state / smoothing seed of this rank
  real* scratch;          // D + D*D
  // smoother exchange only
  real n_obs, d_obs;      // total number of observations n and their dimension d (sigma^2 = sum / n / d)
  int calibrate;
  real* cscale;           // out: sqrt(sigma^2) or 1
  real* scalars;          // out: POF scalars vector (nll, ssq, ssq_proper, cscale slots), or null
};

// register-resident tree sweeps (pof_treelane.cuh), 2D <= 32; nullptr -> CTA-per-node tile kernels
struct TreeLaunch {
  typedef cudaError_t (*Fn)(cudaStream_t, const real* a, long na, const real* b, real* c, long nb);
  Fn fup, fdown, sup, sdown, fcomb, scomb, chunkk, sseed;
  typedef cudaError_t (*FlowFn)(cudaStream_t, const FlowArgs&);
  FlowFn fflow, sflow;     // whole-sweep dataflow kernels (filtering / smoothing)
  typedef cudaError_t (*ExchangeFn)(cudaStream_t, const ExchangeArgs&);
  ExchangeFn fexchange, sexchange;  // rank-carry exchanges of the time-sharded pass (pof_tree_kernels.cuh)
};
const TreeLaunch* tree_launch_a(int D);
const TreeLaunch* tree_launch_b(int D);
const TreeLaunch* tree_launch_c(int D);

}  // namespace POF_NS
