#include "pof_tree_kernels.cuh"
namespace POF_NS {
const TreeLaunch* tree_launch_c(int D) {
  switch (D) {
    case 15: return TreeLaunchers<15>::get();
    case 16: return TreeLaunchers<16>::get();
    default: return nullptr;
  }
}
}
