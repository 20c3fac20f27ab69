"""Large-N helpers for the CPU oracle (TEST INFRASTRUCTURE ONLY, see pof_oracle.py).

The oracle's arithmetic is unchanged: the same `pof_oracle` functions run, only
  * `tria` (pof/utils.py:33-41) splits its batch over a thread pool -- NumPy's batched LAPACK QR is single-threaded
    and releases the GIL, so the split is bitwise identical to the unsplit call (every matrix is factorised alone);
  * the linearisation (pof/step.py:12-22, pof/observations.py:35-40) is evaluated with NumPy-vectorised versions of
    the same SymPy expressions instead of one Python call per time step.
Used by the N >= 2^17 parity tests and by bench.py's CPU legs (which time the reference algorithm at the real N).
"""
from __future__ import annotations

import contextlib
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import sympy as sp

from . import pof_oracle as O

_BASE_TRIA = O.tria


@contextlib.contextmanager
def threads(nthreads=None):
    """`with threads(16): ...` -- inside, O.tria splits batches of >= 4*nthreads matrices over nthreads workers."""
    nthreads = int(nthreads or os.cpu_count() or 1)
    if nthreads <= 1:
        yield O
        return
    pool = ThreadPoolExecutor(max_workers=nthreads)

    def tria(A):
        if A.ndim < 3 or A.shape[0] < 4 * nthreads:
            return _BASE_TRIA(A)
        bounds = np.linspace(0, A.shape[0], nthreads + 1).astype(np.int64)
        outs = list(pool.map(lambda i: _BASE_TRIA(A[bounds[i]:bounds[i + 1]]), range(nthreads)))
        return np.concatenate(outs, axis=0)

    prev = O.tria
    O.tria = tria
    try:
        yield O
    finally:
        O.tria = prev
        pool.shutdown(wait=True)


def _vector_field(ivp):
    """NumPy-vectorised f and Jacobian of a built-in oracle IVP (same SymPy expressions as oracle.ivps._make)."""
    cache = getattr(_vector_field, "_cache", None)
    if cache is None:
        cache = _vector_field._cache = {}
    key = id(ivp.f)
    if key in cache and cache[key][0] is ivp.f:
        return cache[key][1:]
    d = ivp.y0.shape[0]
    ys = sp.symbols(f"y0:{d}")
    fvec = sp.Matrix(list(ivp.exprs))
    jac = fvec.jacobian(sp.Matrix(ys))
    f_l = sp.lambdify(ys, list(fvec), "numpy")
    j_l = sp.lambdify(ys, jac.tolist(), "numpy")
    cache[key] = (ivp.f, f_l, j_l)
    return f_l, j_l


def linearize_at(setup, means):
    """Vectorised oracle.pof_oracle.linearize_at (same formulas): H = E1 - J_f(E0 m) E0, b = (E1 m - f(E0 m)) - H m."""
    ivp, E0, E1 = setup["ivp"], setup["E0"], setup["E1"]
    n = means.shape[0]
    d, D = E0.shape
    f_l, j_l = _vector_field(ivp)
    Y = means @ E0.T  # (n, d)
    cols = [Y[:, i] for i in range(d)]
    fv = np.stack([np.broadcast_to(np.asarray(v, dtype=np.float64), (n,)) for v in f_l(*cols)], axis=1)
    Jrows = j_l(*cols)
    J = np.empty((n, d, d))
    for a in range(d):
        for b in range(d):
            J[:, a, b] = np.broadcast_to(np.asarray(Jrows[a][b], dtype=np.float64), (n,))
    H = E1[None] - J @ E0[None]
    res = means @ E1.T - fv
    b = res - np.einsum("nij,nj->ni", H, means)
    return O.AffineModel(H, b, np.zeros((n, d, d)))


def ieks_step(setup, states, calibrate=True, nthreads=None, scan=O.associative_scan):
    """oracle.pof_oracle.ieks_step (step.py:33-45) with the vectorised linearisation and the threaded `tria`.
    Returns (states, nll, obj, ssq, ssq_proper)."""
    dom = linearize_at(setup, states.mean[1:])
    with threads(nthreads):
        out, nll, obj, ssq, ssqp = O.linear_filtsmooth(setup["x0"], setup["dtm"], dom, scan=scan)
    if calibrate:
        out = O.MVNSqrt(out.mean, np.sqrt(ssq) * out.chol)
    return out, nll, obj, ssq, ssqp
