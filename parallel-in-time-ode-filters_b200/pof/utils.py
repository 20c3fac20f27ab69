"""Types and small helpers mirroring reference pof/utils.py (MVNSqrt :9-11, mvn_loglikelihood :22-30, tria / qr
:33-41, append_zeros_along_new_axis :93-94, objective_function_value :97-101, _gmul :104-107, whiten :110-112) as
plain torch functions on the tensors' device.  The CUDA kernels do not call these: they are the reference's public
helpers, used by its tests and by the baseline paths of this package."""
import math
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


def qr(A):
    """reference utils.py:38-41: the R factor only (LAPACK / cuSOLVER sign convention: diagonal not normalised)"""
    return torch.linalg.qr(A, mode="r").R


def tria(A):
    """reference utils.py:33-35: lower-triangular L with L L^T = A A^T"""
    return qr(A.transpose(-1, -2)).transpose(-1, -2)


def mvn_loglikelihood(x, chol_cov):
    """reference utils.py:22-30: log N(x; 0, chol_cov chol_cov^T)"""
    y = torch.linalg.solve_triangular(chol_cov, x.unsqueeze(-1), upper=False).squeeze(-1)
    dim = chol_cov.shape[-1]
    normalizing_constant = torch.diagonal(chol_cov, dim1=-2, dim2=-1).abs().log().sum(-1) + dim * math.log(2 * math.pi) / 2.0
    return -0.5 * (y * y).sum(-1) - normalizing_constant


def whiten(m, cholP):
    """reference utils.py:110-112 -- solves with cholP^T (upper triangular), as upstream"""
    return torch.linalg.solve_triangular(cholP.transpose(-1, -2), m.unsqueeze(-1), upper=True).squeeze(-1)


def objective_function_value(mnext, m, transition_model):
    """reference utils.py:97-101: || QL^-1 (mnext - F m) ||^2"""
    F, QL = transition_model
    r = torch.linalg.solve_triangular(QL, (mnext - F @ m).unsqueeze(-1), upper=False).squeeze(-1)
    return torch.dot(r, r)


def append_zeros_along_new_axis(z, N):
    """reference utils.py:93-94"""
    return torch.cat([z[None, ...], torch.zeros((N,) + tuple(z.shape), dtype=z.dtype, device=z.device)])
