"""Types and small helpers mirroring reference pof/utils.py (MVNSqrt :9-11, _gmul :104-107)."""
from typing import Any, NamedTuple

import torch


class MVNSqrt(NamedTuple):
    mean: Any
    chol: Any


def _gmul(A, x: MVNSqrt):
    """Multiply a Gaussian with a matrix: A * x (reference pof/utils.py:104-107)."""
    return MVNSqrt(A @ x.mean, A @ x.chol)


def as_f64(x, device=None):
    if isinstance(x, torch.Tensor):
        t = x.to(dtype=torch.float64)
        return t.to(device) if device is not None else t
    return torch.as_tensor(x, dtype=torch.float64, device=device)
