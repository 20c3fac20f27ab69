// The execution context of the C ABI (include/pof_b200.h: pof_ctx_t), shared by the fp64 and fp32 builds of pof_api.cu.
#pragma once
#include <cuda_runtime.h>

#include "../../include/pof_b200.h"

// caller-owned execution context: the side stream on which the smoother's up-sweep runs concurrently with the filter
// scan (fork/join through events, so a pass stays stream-ordered on the caller's stream and is capturable), and the
// optional per-segment timing state.  One context per concurrently running pass; no library-global state.
struct pof_ctx {
  int dev = 0;
  cudaStream_t s2 = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  // ---- per-segment device timing (bench.py): CUDA events recorded on the launching stream around each segment
  static constexpr int MAXP = 4096;
  bool prof_on = false;
  cudaEvent_t ev[MAXP][2];
  int seg[MAXP];
  int created = 0, used = 0;
  double acc[POF_SEG_COUNT] = {0};
  long cnt[POF_SEG_COUNT] = {0};
};


// the IEKS loop as a graph with a WHILE conditional node (pof_ieks_loop_create_*)
struct pof_loop {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
};
