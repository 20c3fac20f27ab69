"""Builds the jax.ffi adapter (pof_b200_ffi.cc -> libpof_b200_ffi.so) IF jaxlib's XLA-FFI headers are discoverable.

    python parallel-in-time-ode-filters_b200/jax_ffi/build.py

The adapter is optional: the product path of this repository (torch tensors through ctypes, pof/_native.py) does not
need it.  In the image this repository was developed in, jax / jaxlib are not installed and the headers
(xla/ffi/api/ffi.h, shipped inside jaxlib) do not exist, so the function below reports "skipped" there.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)


def ffi_include_dir():
    """directory that contains xla/ffi/api/ffi.h, or None"""
    try:
        import jax.ffi  # jax >= 0.4.38

        inc = jax.ffi.include_dir()
    except Exception:
        try:
            import jaxlib

            inc = os.path.join(os.path.dirname(jaxlib.__file__), "include")
        except Exception:
            return None
    return inc if os.path.exists(os.path.join(inc, "xla", "ffi", "api", "ffi.h")) else None


def build(verbose=True):
    inc = ffi_include_dir()
    if inc is None:
        if verbose:
            print("jax_ffi: skipped (no jaxlib / XLA-FFI headers in this environment)")
        return None
    out = os.path.join(HERE, "libpof_b200_ffi.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-std=c++17", "-O2", "-shared", "-Xcompiler", "-fPIC", "-I", inc, os.path.join(HERE, "pof_b200_ffi.cc"),
           "-o", out, "-L", PKG, "-lpof_b200", "-Xlinker", "-rpath", "-Xlinker", PKG]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("jax_ffi build failed:\n" + r.stdout + r.stderr)
    if verbose:
        print("built", out)
    return out


if __name__ == "__main__":
    build()
    sys.exit(0)
