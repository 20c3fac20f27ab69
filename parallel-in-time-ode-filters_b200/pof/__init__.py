"""pof -- parallel-in-time probabilistic ODE filters, B200-native hot path.

Drop-in for the IEKS path of nathanaelbosch/parallel-in-time-ode-filters: `pof.solver.solve`,
`pof.solver.sequential_eks_solve`, `pof.ivp.*`, `pof.step.ieks_step`, `pof.parallel_filtsmooth.linear_filtsmooth`.
Arrays are float64 torch CUDA tensors; all numerics of the path run in libpof_b200.so (CUDA, sm_100a).
"""
from . import _native  # noqa: F401  (raises if the CUDA library is not built: no CPU fallback)
from .utils import MVNSqrt  # noqa: F401
