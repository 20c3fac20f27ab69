"""Oracle-side definitions of the `pof.ivp` problems (reference: pof/ivp.py:7-152).

TEST INFRASTRUCTURE ONLY (see pof_oracle.py).  Vector fields are written symbolically (SymPy) so that
the Jacobian (reference: `jax.jacfwd`, pof/observations.py:38) and the Taylor-mode initial derivatives
(reference: `tornadox.init.TaylorMode`, pof/initialization.py:15-22, un-vendored) are exact and are
derived independently of the product's hand-written CUDA/torch vector fields.
"""
from __future__ import annotations

import functools

import numpy as np
import sympy as sp

from .pof_oracle import IVP


def _make(name, exprs_fn, y0, t0, tmax):
    y0 = np.asarray(y0, dtype=np.float64)
    d = y0.shape[0]
    ys = sp.symbols(f"y0:{d}")
    exprs = [sp.sympify(e) for e in exprs_fn(ys)]
    fvec = sp.Matrix(exprs)
    jac = fvec.jacobian(sp.Matrix(ys))
    f_l = sp.lambdify(ys, list(fvec), "math")
    j_l = sp.lambdify(ys, jac.tolist(), "math")

    def f(t, y):
        return np.array(f_l(*[float(v) for v in y]), dtype=np.float64)

    def jacf(y):
        return np.array(j_l(*[float(v) for v in y]), dtype=np.float64)

    @functools.lru_cache(maxsize=None)
    def taylor(order):
        # y^{(0)} = y ; y^{(k+1)} = (d y^{(k)} / d y) f   (derivatives, not Taylor coefficients)
        cur = sp.Matrix(ys)
        rows = []
        subs = {s: sp.Float(float(v), 40) for s, v in zip(ys, y0)}
        for k in range(order + 1):
            rows.append([float(e.evalf(30, subs=subs)) for e in cur])
            if k < order:
                cur = cur.jacobian(sp.Matrix(ys)) * fvec
        return np.array(rows, dtype=np.float64)

    return IVP(name=name, f=f, jac=jacf, y0=y0, t0=float(t0), tmax=float(tmax), taylor=taylor, exprs=tuple(exprs))


def logistic(t0=0.0, tmax=10.0, y0=None):
    """pof/ivp.py:7-14."""
    y0 = [0.01] if y0 is None else y0
    return _make("logistic", lambda y: [1.0 * y[0] * (1 - y[0])], y0, t0, tmax)


def lotkavolterra(t0=0.0, tmax=7.0, y0=None, p=None):
    """pof/ivp.py:17-31."""
    y0 = [1.0, 1.0] if y0 is None else y0
    a, b, c, dd = p or (1.5, 1.0, 3.0, 1.0)
    return _make(
        "lotkavolterra",
        lambda Y: [a * Y[0] - b * Y[0] * Y[1], -c * Y[1] + dd * Y[0] * Y[1]],
        y0, t0, tmax,
    )


def vanderpol(t0=0.0, tmax=6.3, y0=None, stiffness_constant=1e1):
    """pof/ivp.py:34-41."""
    y0 = [2.0, 0.0] if y0 is None else y0
    mu = stiffness_constant
    return _make("vanderpol", lambda Y: [Y[1], mu * ((1.0 - Y[0] ** 2) * Y[1] - Y[0])], y0, t0, tmax)


def fitzhughnagumo(t0=0.0, tmax=100.0, y0=None, p=None):
    """pof/ivp.py:44-60."""
    y0 = [1.0, 1.0] if y0 is None else y0
    a, b, tinv, l = p or (0.7, 0.8, 1 / 12.5, 0.5)
    return _make(
        "fitzhughnagumo",
        lambda Y: [Y[0] - (Y[0] ** 3) / 3 - Y[1] + l, tinv * (Y[0] + a - b * Y[1])],
        y0, t0, tmax,
    )


def rober(t0=0.0, tmax=1e11, y0=None, p=None):
    """pof/ivp.py:63-79."""
    y0 = [1.0, 0.0, 0.0] if y0 is None else y0
    k1, k2, k3 = p or (0.04, 3e7, 1e4)
    return _make(
        "rober",
        lambda y: [-k1 * y[0] + k3 * y[1] * y[2], k1 * y[0] - k2 * y[1] ** 2 - k3 * y[1] * y[2], k2 * y[1] ** 2],
        y0, t0, tmax,
    )


def rigid_body(t0=0.0, tmax=20.0, y0=None, p=None):
    """pof/ivp.py:82-90."""
    y0 = [1.0, 0.0, 0.9] if y0 is None else y0
    p = p or (-2.0, 1.25, -0.5)
    return _make(
        "rigid_body", lambda y: [p[0] * y[1] * y[2], p[1] * y[0] * y[2], p[2] * y[0] * y[1]], y0, t0, tmax
    )


def seir(t0=0.0, tmax=200.0, y0=None, p=None):
    """pof/ivp.py:93-109."""
    y0 = [998.0, 1.0, 1.0, 1.0] if y0 is None else y0
    p = p or (0.3, 0.3, 0.1, float(np.sum(y0)))
    return _make(
        "seir",
        lambda y: [
            -p[1] * y[0] * y[2] / p[3],
            p[1] * y[0] * y[2] / p[3] - p[0] * y[1],
            p[0] * y[1] - p[2] * y[2],
            p[2] * y[2],
        ],
        y0, t0, tmax,
    )


def threebody(t0=0.0, tmax=17.0652165601579625588917206249, y0=None, p=None):
    """pof/ivp.py:112-126."""
    y0 = [0.994, 0.0, 0.0, -2.00158510637908252240537862224] if y0 is None else y0
    mu = (p or (0.012277471,))[0]
    mp = 1.0 - mu

    def ex(y):
        D1 = sp.sqrt((y[0] + mu) ** 2 + y[1] ** 2) ** 3
        D2 = sp.sqrt((y[0] - mp) ** 2 + y[1] ** 2) ** 3
        du0p = y[0] + 2 * y[3] - mp * (y[0] + mu) / D1 - mu * (y[0] - mp) / D2
        du1p = y[1] - 2 * y[2] - mp * y[1] / D1 - mu * y[1] / D2
        return [y[2], y[3], du0p, du1p]

    return _make("threebody", ex, y0, t0, tmax)


def henonheiles(t0=0.0, tmax=100.0, y0=None, p=None):
    """pof/ivp.py:137-152."""
    y0 = [0.5, 0.0, 0.0, 0.1] if y0 is None else y0
    p = 1.0 if p is None else p
    return _make(
        "henonheiles",
        lambda y: [y[2], y[3], -y[0] - 2 * p * y[0] * y[1], -y[1] - p * (y[0] ** 2 - y[1] ** 2)],
        y0, t0, tmax,
    )


def lorenz96(t0=0.0, tmax=10.0, y0=None, d=16, forcing=8.0):
    """Not in pof/ivp.py: the larger-state problem of BASELINE config 5: Lorenz-96,
    f_i = (y_{i+1} - y_{i-2}) y_{i-1} - y_i + F (cyclic), y0_i = F + 0.01 sin(i)."""
    if y0 is None:
        y0 = [forcing + 0.01 * np.sin(i) for i in range(d)]
    d = len(y0)
    return _make(
        "lorenz96",
        lambda y: [(y[(i + 1) % d] - y[(i - 2) % d]) * y[(i - 1) % d] - y[i] + forcing for i in range(d)],
        y0, t0, tmax,
    )


ALL = dict(
    logistic=logistic, lotkavolterra=lotkavolterra, vanderpol=vanderpol, fitzhughnagumo=fitzhughnagumo,
    rober=rober, rigid_body=rigid_body, seir=seir, threebody=threebody, henonheiles=henonheiles,
    lorenz96=lorenz96,
)
