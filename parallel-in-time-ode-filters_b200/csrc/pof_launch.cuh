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
  const real* stop;  // device-side IEKS loop (pof_ieks_loop_step): if non-null and *stop != 0 the loop has ended and
                     // every kernel that touches persistent state (or does real work) returns at once
  int no_tma;        // lane2 smoother: 0 = bulk-copy (TMA) staging of the backward kernels (POF_F_SMOOTH_TMA; measured slower)
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
  enum Kind { UP = 0, ROOT = 1, DOWN = 2, DOWN_E = 3, KS = 4, KS_APPLY = 5 };
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
  const real* stop;         // as LeafArgs::stop
  // ---- hybrid filter sweep: the levels above `ks_base` are replaced by ONE Kogge-Stone inclusive scan over the ks_n
  // nodes of that level (ks_steps = ceil(log2 ks_n) dependent steps instead of 2 x that many up/down levels; the extra
  // combines run on warps that the narrow top levels leave idle anyway).  Step s (KS segment, level = s) combines, for
  // every node j >= 2^(s-1), the window ending at j - 2^(s-1) with the node's own window; a node's window element is
  // final after step nbits(j) and is never copied: step s reads node j' at level min(s-1, nbits(j')).  Level 0 = the
  // elements of tree level ks_base (agg), level s >= 1 = ks + (s-1) * ks_n elements.  KS_APPLY then turns the
  // inclusive elements into incoming states: state(j) = root state (+) element(j-1), and the down-sweep continues
  // from level ks_base.
  int ks_base, ks_steps;
  long ks_n;
  real* ks;
  unsigned* flag_ks;        // ks_steps x ks_n, zeroed with the other flags
  int ks_wait;              // 1: the KS elements are built by this launch (wait for their flags); 0: complete before
  int poll_ns;              // back-off between two polls of the dependency flags
};

struct ExchangeArgs {
  int rank, world;
  const real* gathered;   // world payloads, `stride` doubles each (ignored with peer memory: the local slots are used)
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
  // ---- peer-memory form (pof_p2p_*): the kernel ITSELF moves the payload -- it stores this rank's payload into its
  // slot of every peer's exchange area over NVLink, releases a flag there, waits for the flags of the ranks it
  // needs in its own area, and folds: compute and collective in one launch, no NCCL call.  p2p = 0: plain form.
  int p2p;
  const real* payload;          // this rank's payload (stride values)
  unsigned long long* peer[8];  // exchange areas of all ranks (8-byte words; peer[rank] is the local one)
  long epoch_word, flag_word, slot_word;  // word offsets: this exchange's epoch counter, flags[world], slots[2][world][stride]
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Peer-memory exchange prologue (one warp): push `payload` into this rank's slot of every rank's exchange area, release
// the flag there, wait for the ranks < need_below ... i.e. all r with need(r).  Flags carry a monotonically increasing
// epoch (kept in the local area, advanced by this kernel), slots are double-buffered by its parity.  A wait that
// exceeds ~2 s (a peer that died) sets the area's status word instead of hanging the GPU.
// Returns the address of the gathered payloads (world x stride) in the LOCAL area.
__device__ __forceinline__ const real* p2p_exchange(const ExchangeArgs& A, int need_lo, int need_hi) {
  const int lane = threadIdx.x & 31;
  const int W = A.world;
  unsigned long long* mine = A.peer[A.rank];
  const unsigned long long e = mine[A.epoch_word] + 1ull;
  const long par = (long)(e & 1ull);
  const long slot = (par * W + A.rank) * A.stride;  // (in values, inside the slot region that starts at slot_word)
  for (int p = 0; p < W; ++p) {
    real* dst = reinterpret_cast<real*>(A.peer[p] + A.slot_word) + slot;
    for (long j = lane; j < A.stride; j += 32) dst[j] = A.payload[j];
  }
  __threadfence_system();
  __syncwarp();
  if (lane < W) st_release_sys(A.peer[lane] + A.flag_word + A.rank, e);
  const bool need = lane < W && lane >= need_lo && lane < need_hi && lane != A.rank;
  const long long t0 = clock64();
  while (true) {
    const bool ok = !need || ld_acquire_sys(mine + A.flag_word + lane) >= e;
    if (__all_sync(0xffffffffu, ok)) break;
    if (clock64() - t0 > 4000000000ll) {
      if (lane == 0) mine[0] = 1ull;  // status: timed out
      break;
    }
    __nanosleep(100);
  }
  __syncwarp();
  if (lane == 0) mine[A.epoch_word] = e;
  return reinterpret_cast<const real*>(mine + A.slot_word) + par * W * A.stride;
}

#endif  // __CUDACC__


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
