"""Initial state x0 and initial linearisation trajectories (reference pof/initialization.py).

`taylor_mode_init` replaces the un-vendored `tornadox.init.TaylorMode` (initialization.py:15-22) by nested
forward-mode jvps: y^(0) = y0, y^(k+1) = d/dt y^(k) = J_{y^(k)}(y) f(y).  One-off host-side set-up (tiny tensors).
"""
import numpy as np
import torch

from .transitions import IWP
from .utils import MVNSqrt


def _cpu64(y):
    return torch.as_tensor(y, dtype=torch.float64).detach().cpu()


def taylor_derivatives(f, y0, num_derivatives):
    """rows y^{(k)}(t0), k = 0..q, shape (q+1, d)"""
    y0 = _cpu64(y0)

    def fy(y):
        return f(None, y)

    def deriv(k):
        if k == 0:
            return lambda y: y
        prev = deriv(k - 1)
        return lambda y: torch.func.jvp(prev, (y,), (fy(y),))[1]

    rows = [deriv(k)(y0) for k in range(num_derivatives + 1)]
    return torch.stack(rows)


def taylor_mode_init(f, y0, num_derivatives):
    """reference initialization.py:15-22: mean = [y1, y1', .., y1^(q), y2, ..], chol = 0"""
    derivs = taylor_derivatives(f, y0, num_derivatives)  # (q+1, d)
    m0 = derivs.T.reshape(-1).contiguous()
    D = m0.shape[0]
    return MVNSqrt(m0, torch.zeros((D, D), dtype=torch.float64))


def constant_init(*, y0, order, ts, f=True):
    """reference initialization.py:42-56"""
    y0 = _cpu64(y0)
    d = y0.shape[-1]
    N = len(ts)
    dy0 = _cpu64(f(None, y0)) if f is not None else torch.zeros_like(y0)
    x0 = torch.cat([y0[:, None], dy0[:, None], torch.zeros((d, order - 1), dtype=torch.float64)], dim=1).reshape(1, -1)
    traj = x0.repeat(N, 1)
    D = traj.shape[1]
    return MVNSqrt(traj, torch.zeros((N, D, D), dtype=torch.float64))


def prior_init(*, f, y0, order, ts, device=None, means_only=False, x0=None):
    """reference initialization.py:66-89, including its quirk: the step sizes are the absolute times ts[1:], and the
    trajectory is returned in NON-preconditioned coordinates.  Row k is closed form (one prediction of x0 = (m0, 0)),
    evaluated by the CUDA kernel `pof_prior_init_f64` (one thread per row entry); only the means feed the first
    linearisation, so `solve` asks for `means_only`."""
    import ctypes  # noqa: F401

    from . import _native as nat
    from .transitions import preconditioned_discretize_1d

    if x0 is None:  # (set_up_solver has computed it already: `setup["_x0_raw"]`)
        x0 = taylor_mode_init(f, y0, order)
    d = int(_cpu64(y0).shape[0])
    iwp = IWP(num_derivatives=order, wiener_process_dimension=d)
    _, qL = preconditioned_discretize_1d(iwp)
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    ts_d = torch.as_tensor(np.asarray(_cpu64(ts)), dtype=torch.float64).to(dev).contiguous()
    N = ts_d.shape[0]
    D = x0.mean.shape[0]
    m0 = x0.mean.to(dev).contiguous()
    means = torch.empty((N, D), dtype=torch.float64, device=dev)
    chols = None if means_only else torch.empty((N, D, D), dtype=torch.float64, device=dev)
    qLh, qLp = nat.host_doubles(qL)
    nat.require_cuda(ts_d, m0, means, chols)
    nat.check(nat.LIB.pof_prior_init_f64(nat.stream_ptr(), N, d, order, qLp, nat.ptr(ts_d), nat.ptr(m0),
                                         nat.ptr(means), nat.ptr(chols)), "pof_prior_init_f64")
    return MVNSqrt(means, chols)


def coarse_ekf_init(*, f, y0, order, ts, N=10):
    """reference initialization.py:103-121: sequential EKS (GPU kernel `pof_sequential_eks_f64`) on a coarse grid of N
    points with full states, then piecewise-constant interpolation idx = floor(ts / coarse_dt) (absolute times: the
    reference assumes t0 = 0; out-of-range indices clamp as in JAX)."""
    from .solver import sequential_eks_solve

    ts = np.asarray(_cpu64(ts))
    coarse_ts = np.linspace(ts[0], ts[-1], N)
    coarse_dt = coarse_ts[1] - coarse_ts[0]
    out, _ = sequential_eks_solve(f=f, y0=y0, ts=coarse_ts, order=order, return_full_states=True)
    idxs = np.clip(np.floor(ts / coarse_dt).astype(np.int64), 0, N - 1)
    idx_t = torch.from_numpy(idxs).to(out.mean.device)
    return MVNSqrt(out.mean.index_select(0, idx_t), out.chol.index_select(0, idx_t))


def _prior_init(*, x0, dtm):
    """reference initialization.py:75-80: every row k >= 1 is x0 pushed through transition k ALONE (not the product of
    the transitions up to k): mean F_k m0, factor tria([F_k L0 | QL_k]); row 0 is x0.  dtm: (n,D,D) models."""
    from .utils import MVNSqrt, tria

    F, QL = dtm.F, dtm.QL
    if F.dim() != 3:
        raise ValueError("_prior_init needs per-step transition models (n,D,D), e.g. from discretize_transitions")
    n = F.shape[0]
    means = torch.einsum("nij,j->ni", F, x0.mean)
    chols = tria(torch.cat([F @ x0.chol.unsqueeze(0).expand(n, *x0.chol.shape), QL], dim=-1))
    return MVNSqrt(torch.cat([x0.mean[None], means]), torch.cat([x0.chol[None], chols]))


def updated_prior_init(*, x0, dtm, om):
    """reference initialization.py:92-100: the prior trajectory with every state updated on its own linearised
    (noiseless) observation"""
    from .observations import linearize
    from .sequential_filtsmooth.loops import _sqrt_update
    from .utils import MVNSqrt

    states = _prior_init(x0=x0, dtm=dtm)
    out = []
    for k in range(states.mean.shape[0]):
        x = MVNSqrt(states.mean[k], states.chol[k])
        H, b, cholR = linearize(om, x)
        out.append(_sqrt_update(H, cholR, b, x)[0])
    return MVNSqrt(torch.stack([o.mean for o in out]), torch.stack([o.chol for o in out]))


def uncertain_init(f, y0, num_derivatives, var=1.0):
    """reference initialization.py:25-40: mean [y0, f(y0), 0, ...] per dimension, unit-variance (x var) factors on the
    derivatives of order >= 2, zero variance on y0 and f(y0)"""
    from .utils import MVNSqrt

    y0 = _cpu64(y0)
    d, q = y0.shape[0], num_derivatives
    dy0 = _cpu64(f(None, y0))
    m0 = torch.cat([y0[:, None], dy0[:, None], torch.zeros((d, q - 1), dtype=torch.float64)], dim=1).reshape(-1)
    diag = torch.full((d * (q + 1),), float(var), dtype=torch.float64)
    for j in range(d):
        diag[(q + 1) * j] = 0.0
        diag[(q + 1) * j + 1] = 0.0
    return MVNSqrt(m0, torch.diag(diag.sqrt()))


def classic_to_init(*, ys, order, f=None):
    """reference initialization.py:58-72: a trajectory of ODE states ys (N, d) -> state trajectory
    [y, f(y), 0, ...] per dimension with zero covariance factors"""
    from .utils import MVNSqrt

    ys = _cpu64(ys)
    N, d = ys.shape
    dys = torch.stack([_cpu64(f(None, y)) for y in ys]) if f is not None else torch.zeros_like(ys)
    traj = torch.cat([ys[:, :, None], dys[:, :, None], torch.zeros((N, d, order - 1), dtype=torch.float64)], dim=2)
    traj = traj.reshape(N, -1)
    D = traj.shape[1]
    return MVNSqrt(traj, torch.zeros((N, D, D), dtype=torch.float64))
