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
