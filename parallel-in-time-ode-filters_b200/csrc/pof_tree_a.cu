#include "pof_tree_kernels.cuh"
namespace POF_NS {
const TreeLaunch* tree_launch_a(int D) {
  switch (D) {
    case 2: return TreeLaunchers<2>::get();
    case 3: return TreeLaunchers<3>::get();
    case 4: return TreeLaunchers<4>::get();
    case 5: return TreeLaunchers<5>::get();
    case 6: return TreeLaunchers<6>::get();
    case 8: return TreeLaunchers<8>::get();
    default: return nullptr;
  }
}
}
