"""One IEKS iteration (reference pof/step.py:12-45)."""
import torch

from . import _native as nat
from .observations import AffineModel
from .parallel_filtsmooth import linear_filtsmooth as parallel_linear_filtsmooth
from .sequential_filtsmooth import linear_filtsmooth as sequential_linear_filtsmooth
from .utils import MVNSqrt


def linearize_into(lin, means, H, c):
    """H[k], c[k] <- linearisation of x -> E1 x - f(E0 x) at means[k+1] (reference step.py:12-22,
    observations.py:35-40).  Built-in vector fields: fused CUDA kernel; user f: torch.func autodiff on the device."""
    n = means.shape[0] - 1
    d, q = lin["d"], lin["q"]
    if lin["builtin"] is not None:
        ivp_id, params = lin["builtin"]
        ph, pp = nat.host_doubles(list(params) + [0.0])
        nat.require_cuda(means, H, c)
        rc = nat.LIB.pof_linearize_ivp_f64(nat.stream_ptr(), ivp_id, pp, len(params), n, d, q, lin["scale0"],
                                           lin["scale1"], nat.ptr(means[1:]), nat.ptr(H), nat.ptr(c))
        nat.check(rc, "pof_linearize_ivp_f64")
        return
    f, E0, E1 = lin["f"], lin["E0"], lin["E1"]
    ys = means[1:] @ E0.T  # (n, d)
    fy = lambda y: f(None, y)
    J = torch.func.vmap(torch.func.jacfwd(fy))(ys)  # (n, d, d)
    fv = torch.func.vmap(fy)(ys)
    H.copy_(E1.unsqueeze(0) - J @ E0.unsqueeze(0))
    c.copy_(torch.einsum("nij,nj->ni", J, ys) - fv)


def linearize_at_previous_states(om, prev_states):
    """reference step.py:12-22 -> AffineModel(H (n,d,D), b (n,d), cholR (n,d,d) = 0)"""
    lin = om.f._pof_lin
    means = prev_states.mean.contiguous()
    n, D = means.shape[0] - 1, means.shape[1]
    d = lin["d"]
    H = torch.empty((n, d, D), dtype=torch.float64, device=means.device)
    c = torch.empty((n, d), dtype=torch.float64, device=means.device)
    linearize_into(lin, means, H, c)
    return AffineModel(H, c, torch.zeros((n, d, d), dtype=torch.float64, device=means.device))


def inflate(state):
    """reference step.py:26-30: add 1e-3 to the diagonal of a covariance factor"""
    eye = torch.eye(state.chol.shape[0], dtype=state.chol.dtype, device=state.chol.device)
    return MVNSqrt(state.mean, state.chol + eye * 1e-3)


def ieks_step(*, om, dtm, x0, states, calibrate=True, sequential=False):
    """reference step.py:33-45"""
    dom = linearize_at_previous_states(om, states)
    if not sequential:
        states, nll, obj, ssq = parallel_linear_filtsmooth(x0, dtm, dom)
    else:
        states, nll, obj, ssq = sequential_linear_filtsmooth(x0, dtm, dom)
    if calibrate:
        states = MVNSqrt(states.mean, ssq**0.5 * states.chol)
    return states, nll, obj, ssq
