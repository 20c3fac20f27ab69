"""JAX side of the jax.ffi adapter: registers the handlers of libpof_b200_ffi.so and wraps them with the reference's
function signatures (`pof.parallel_filtsmooth.linear_filtsmooth`, the body of `pof.step.ieks_step`).

This is the stub a maintainer of the reference adds.  It needs jax + jaxlib with CUDA and the
adapter built by jax_ffi/build.py; neither exists in the image this repository was developed in, so this module is NOT
imported by the package or by any test -- it is untested reference material kept next to the C++ source it binds.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    import jax

    lib = ctypes.CDLL(os.path.join(HERE, "libpof_b200_ffi.so"))
    core = ctypes.CDLL(os.path.join(os.path.dirname(HERE), "libpof_b200.so"))
    core.pof_workspace_bytes.restype = ctypes.c_size_t
    core.pof_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int64]
    core.pof_default_chunk_len.restype = ctypes.c_int64
    core.pof_default_chunk_len.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint32]
    for name, sym in [("pof_filtsmooth", "PofFiltSmooth"), ("pof_ieks_iteration", "PofIeksIteration"),
                      ("pof_linearize", "PofLinearize"), ("pof_filter_combine", "PofFilterCombine"),
                      ("pof_smooth_combine", "PofSmoothCombine"), ("pof_project", "PofProject")]:
        jax.ffi.register_ffi_target(name, jax.ffi.pycapsule(getattr(lib, sym)), platform="CUDA")
    return core


def linear_filtsmooth(x0, dtm, dom, *, qL, sm_count=148):
    """reference parallel_filtsmooth/__init__.py:5-10 for the preconditioned IWP prior with noiseless observations.
    x0 = (mean (D,), chol (D,D)); dom = (H (n,d,D), b (n,d), cholR == 0); qL: the (q+1,q+1) block of dtm.QL (host)."""
    import jax
    import jax.numpy as jnp

    core = _load()
    n, d, D = dom.H.shape
    N, q = n + 1, D // d - 1
    L = int(core.pof_default_chunk_len(N, d, q, sm_count, 0))
    ws_bytes = int(core.pof_workspace_bytes(N, d, q, L))
    f64 = jnp.float64
    means, chols, scal, _ = jax.ffi.ffi_call(
        "pof_filtsmooth",
        (jax.ShapeDtypeStruct((N, D), f64), jax.ShapeDtypeStruct((N, D, D), f64), jax.ShapeDtypeStruct((8,), f64),
         jax.ShapeDtypeStruct((ws_bytes,), jnp.uint8)),
    )(x0.mean, x0.chol, dom.H, dom.b, jnp.zeros((N, D), f64), chunk_len=np.int64(L), calibrate=False,
      flags=np.int64(0), qL=np.asarray(qL, dtype=np.float64).ravel())
    return (means, chols), scal[0], scal[1], scal[2]  # (states), nll, obj, ssq   (POF_S_* indices)
