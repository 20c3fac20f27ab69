// fp32 build of pof_tree_b.cu (namespace pof32, C ABI entry points *_f32): see pof_real.cuh
#define POF_F32 1
#include "pof_tree_b.cu"
