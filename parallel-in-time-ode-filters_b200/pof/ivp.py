"""Initial value problems, same factories / defaults as reference pof/ivp.py.

Each factory returns an `InitialValueProblem` with `.f(t, y)`, `.y0`, `.t0`, `.tmax`, `.t_span` (the attributes the
reference's callers use; reference objects are `tornadox.ivp.InitialValueProblem`, un-vendored).  `f` is written with
torch ops (there is no JAX in this environment) and carries a `_pof_builtin = (ivp_id, params)` tag so that the
solver evaluates the vector field and its Jacobian inside the fused CUDA linearisation kernel
(pof_linearize_ivp_f64) instead of through autodiff.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable

import torch

from ._native import IVP_IDS


@dataclass(frozen=True)
class InitialValueProblem:
    f: Callable
    t0: float
    tmax: float
    y0: torch.Tensor

    @property
    def t_span(self):
        return (self.t0, self.tmax)

    @property
    def dimension(self):
        return int(self.y0.shape[0])


def _t(x):
    return torch.as_tensor(x, dtype=torch.float64)


def _tag(f, name, params):
    f._pof_builtin = (IVP_IDS[name], tuple(float(p) for p in params))
    return f


def logistic(t0=0, tmax=10.0, y0=None):
    """reference ivp.py:7-14"""
    y0 = _t([0.01]) if y0 is None else _t(y0)

    def f(t, y):
        return 1.0 * y * (1 - y)

    return InitialValueProblem(f=_tag(f, "logistic", ()), t0=t0, tmax=tmax, y0=y0)


def lotkavolterra(t0=0.0, tmax=7, y0=None, p=None):
    """reference ivp.py:17-31"""
    y0 = _t([1.0, 1.0]) if y0 is None else _t(y0)
    p = (1.5, 1.0, 3.0, 1.0) if p is None else tuple(float(v) for v in p)

    def f(_, Y, p=p):
        a, b, c, d = p
        return torch.stack([a * Y[0] - b * Y[0] * Y[1], -c * Y[1] + d * Y[0] * Y[1]])

    return InitialValueProblem(f=_tag(f, "lotkavolterra", p), t0=t0, tmax=tmax, y0=y0)


def vanderpol(t0=0.0, tmax=6.3, y0=None, stiffness_constant=1e1):
    """reference ivp.py:34-41"""
    y0 = _t([2.0, 0.0]) if y0 is None else _t(y0)
    mu = float(stiffness_constant)

    def f_vanderpol(_, Y, mu=mu):
        return torch.stack([Y[1], mu * ((1.0 - Y[0] ** 2) * Y[1] - Y[0])])

    return InitialValueProblem(f=_tag(f_vanderpol, "vanderpol", (mu,)), t0=t0, tmax=tmax, y0=y0)


def fitzhughnagumo(t0=0.0, tmax=100.0, y0=None, p=None):
    """reference ivp.py:44-60"""
    y0 = _t([1.0, 1.0]) if y0 is None else _t(y0)
    p = (0.7, 0.8, 1 / 12.5, 0.5) if p is None else tuple(float(v) for v in p)

    def f(_, Y, p=p):
        a, b, tinv, l = p
        v = Y[0]
        w = Y[1]
        return torch.stack([v - (v**3) / 3 - w + l, tinv * (v + a - b * w)])

    return InitialValueProblem(f=_tag(f, "fitzhughnagumo", p), t0=t0, tmax=tmax, y0=y0)


def rober(t0=0.0, tmax=1e11, y0=None, p=None):
    """reference ivp.py:63-79"""
    y0 = _t([1.0, 0.0, 0.0]) if y0 is None else _t(y0)
    p = (0.04, 3e7, 1e4) if p is None else tuple(float(v) for v in p)

    def f(_, Y, p=p):
        k1, k2, k3 = p
        y1, y2, y3 = Y[0], Y[1], Y[2]
        return torch.stack([-k1 * y1 + k3 * y2 * y3, k1 * y1 - k2 * y2**2 - k3 * y2 * y3, k2 * y2**2])

    return InitialValueProblem(f=_tag(f, "rober", p), t0=t0, tmax=tmax, y0=y0)


def rigid_body(t0=0.0, tmax=20.0, y0=None, p=None):
    """reference ivp.py:82-90"""
    y0 = _t([1.0, 0.0, 0.9]) if y0 is None else _t(y0)
    p = (-2.0, 1.25, -0.5) if p is None else tuple(float(v) for v in p)

    def f(_, y, p=p):
        return torch.stack([p[0] * y[1] * y[2], p[1] * y[0] * y[2], p[2] * y[0] * y[1]])

    return InitialValueProblem(f=_tag(f, "rigid_body", p), t0=t0, tmax=tmax, y0=y0)


def seir(t0=0.0, tmax=200.0, y0=None, p=None):
    """reference ivp.py:93-109"""
    y0 = _t([998.0, 1.0, 1.0, 1.0]) if y0 is None else _t(y0)
    p = (0.3, 0.3, 0.1, float(y0.sum())) if p is None else tuple(float(v) for v in p)

    def f(_, y, p=p):
        return torch.stack(
            [
                -p[1] * y[0] * y[2] / p[3],
                p[1] * y[0] * y[2] / p[3] - p[0] * y[1],
                p[0] * y[1] - p[2] * y[2],
                p[2] * y[2],
            ]
        )

    return InitialValueProblem(f=_tag(f, "seir", p), t0=t0, tmax=tmax, y0=y0)


def threebody(t0=0.0, tmax=17.0652165601579625588917206249, y0=None, p=None):
    """reference ivp.py:112-126"""
    y0 = _t([0.994, 0.0, 0.0, -2.00158510637908252240537862224]) if y0 is None else _t(y0)
    p = (0.012277471,) if p is None else tuple(float(v) for v in p)

    def f(_, y, p=p):
        mu, mp = p[0], 1.0 - p[0]
        D1 = torch.sqrt((y[0] + mu) ** 2 + y[1] ** 2) ** 3.0
        D2 = torch.sqrt((y[0] - mp) ** 2 + y[1] ** 2) ** 3.0
        du0p = y[0] + 2 * y[3] - mp * (y[0] + mu) / D1 - mu * (y[0] - mp) / D2
        du1p = y[1] - 2 * y[2] - mp * y[1] / D1 - mu * y[1] / D2
        return torch.stack([y[2], y[3], du0p, du1p])

    return InitialValueProblem(f=_tag(f, "threebody", p), t0=t0, tmax=tmax, y0=y0)


def henonheiles(t0=0.0, tmax=100.0, y0=None, p=None):
    """reference ivp.py:137-152"""
    y0 = _t([0.5, 0.0, 0.0, 0.1]) if y0 is None else _t(y0)
    p = 1.0 if p is None else float(p)

    def f(_, y, p=p):
        return torch.stack(
            [y[2], y[3], -y[0] - 2 * p * y[0] * y[1], -y[1] - p * (y[0] ** 2 - y[1] ** 2)]
        )

    return InitialValueProblem(f=_tag(f, "henonheiles", (p,)), t0=t0, tmax=tmax, y0=y0)


def lorenz96(t0=0.0, tmax=10.0, y0=None, d=16, forcing=8.0):
    """Not in the reference's ivp.py: the larger-state problem of BASELINE config 5.  Lorenz-96,
    f_i = (y_{i+1} - y_{i-2}) y_{i-1} - y_i + F (cyclic indices), y0_i = F + 0.01 sin(i); d >= 4."""
    if y0 is None:
        y0 = forcing + 0.01 * torch.sin(torch.arange(d, dtype=torch.float64))
    y0 = _t(y0)
    forcing = float(forcing)

    def f(_, y, forcing=forcing):
        return (torch.roll(y, -1, -1) - torch.roll(y, 2, -1)) * torch.roll(y, 1, -1) - y + forcing

    return InitialValueProblem(f=_tag(f, "lorenz96", (forcing,)), t0=t0, tmax=tmax, y0=y0)
