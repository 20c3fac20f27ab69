#include "pof_tree_kernels.cuh"
namespace POF_NS {
const TreeLaunch* tree_launch_b(int D) {
  switch (D) {
    case 9: return TreeLaunchers<9>::get();
    case 10: return TreeLaunchers<10>::get();
    case 12: return TreeLaunchers<12>::get();
    default: return nullptr;
  }
}
}
