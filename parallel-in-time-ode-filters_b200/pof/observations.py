"""Observation models (reference pof/observations.py:14-40).  Noiseless: cholR = 0."""
from typing import Callable, NamedTuple

import torch


class NonlinearModel(NamedTuple):
    """Nonlinear noiseless observation model: y = f(x) (reference observations.py:14-20)"""

    f: Callable

    def __call__(self, x):
        return self.f(x)


class AffineModel(NamedTuple):
    """Affine approximation y = H x + b of a nonlinear model (reference observations.py:23-33)"""

    H: torch.Tensor
    b: torch.Tensor
    cholR: torch.Tensor


def linearize(f, x):
    """EK1 linearisation at the mean of x (reference observations.py:35-40); autodiff path for user-supplied models."""
    m = x.mean
    res = f(m)
    F_x = torch.func.jacfwd(f)(m)
    cholR = torch.zeros((res.shape[0], res.shape[0]), dtype=m.dtype, device=m.device)
    return AffineModel(F_x, res - F_x @ m, cholR)


def linearize_ek0(f, x):
    """reference observations.py:52-63: zeroth-order linearisation, H = E1 (no Jacobian of the vector field)"""
    m = x.mean
    res = f(m)
    d = res.shape[0]
    q = m.shape[0] // d - 1
    e1 = torch.zeros((1, q + 1), dtype=m.dtype, device=m.device)
    e1[0, 1] = 1.0
    E1 = torch.kron(torch.eye(d, dtype=m.dtype, device=m.device), e1)
    return AffineModel(E1, res - E1 @ m, torch.zeros((d, d), dtype=m.dtype, device=m.device))


def uncertain_linearize(f, x):
    """reference observations.py:43-49 (marked WIP upstream): cholR = tria(F_m chol)"""
    m, CL = x.mean, x.chol
    res = f(m)
    F_m = torch.func.jacfwd(f)(m)
    cholR = torch.linalg.qr((F_m @ CL).T, mode="r").R.T
    return AffineModel(F_m, res - F_m @ m, cholR)


def linearize_regularized(f, x, l):
    """reference observations.py:66-83 (Levenberg-Marquardt regularisation): the EK1 model stacked with a pseudo-
    observation x ~ N(m, I / l); observation dimension d + D.  Kept with the reference's exact (work-in-progress)
    offset convention `full_b = [res, -m]`.  Host-/autodiff-side helper: the CUDA pass takes observations of dimension d."""
    m = x.mean
    res = f(m)
    F_x = torch.func.jacfwd(f)(m)
    d, D = res.shape[0], m.shape[0]
    z = lambda *s: torch.zeros(*s, dtype=m.dtype, device=m.device)
    eye = torch.eye(D, dtype=m.dtype, device=m.device)
    full_b = torch.cat([res, -m])
    full_H = torch.cat([F_x, eye], dim=0)
    full_cholR = torch.cat([torch.cat([z(d, d), z(d, D)], dim=1), torch.cat([z(D, d), eye / l ** 0.5], dim=1)], dim=0)
    return AffineModel(full_H, full_b, full_cholR)
