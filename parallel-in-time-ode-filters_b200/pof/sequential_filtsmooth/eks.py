"""Sequential extended Kalman smoother (reference pof/sequential_filtsmooth/__init__.py:5-10): the EKF is relinearised
at the predicted mean of every step, so the recursion is inherently sequential -- one GPU thread walks the grid
(`pof_sequential_eks_f64`).  Baseline / cross-check path of the reference, kept for API completeness."""
import ctypes

import torch

from .. import _native as nat
from ..utils import MVNSqrt


def eks_filtsmooth(setup):
    """-> (MVNSqrt(means (N,D), chols (N,D,D)), ell, obj, ssq) like `filtsmooth(x0, dtm, om)`"""
    lin = setup["om"].f._pof_lin
    if lin["builtin"] is None:
        return _eks_filtsmooth_user_f(setup)
    d, q = lin["d"], lin["q"]
    D = d * (q + 1)
    N = len(setup["ts"])
    dev = setup["_device"]
    x0 = setup["x0"]
    means = torch.empty((N, D), dtype=torch.float64, device=dev)
    chols = torch.empty((N, D, D), dtype=torch.float64, device=dev)
    scalars = torch.zeros(nat.NSCALARS, dtype=torch.float64, device=dev)
    ws = nat.Workspace.get(N, d, q, N - 1, dev)
    ivp_id, params = lin["builtin"]
    ph, pp = nat.host_doubles(list(params) + [0.0])
    qLh, qLp = nat.host_doubles(setup["_qL"])
    rc = nat.LIB.pof_sequential_eks_f64(
        nat.stream_ptr(), nat.flags(), ivp_id, pp, len(params), N, d, q, qLp, lin["scale0"], lin["scale1"],
        nat.ptr(x0.mean), nat.ptr(x0.chol), nat.ptr(means), nat.ptr(chols), nat.ptr(scalars), ws.ws_ptr, ws.nbytes)
    nat.check(rc, "pof_sequential_eks_f64")
    sc = scalars.cpu()
    return MVNSqrt(means, chols), float(sc[nat.S_NLL]), float(sc[nat.S_OBJ]), float(sc[nat.S_SSQ])


def _tria(A):
    return torch.linalg.qr(A.T, mode="r").R.T


def _eks_filtsmooth_user_f(setup):
    """The same extended Kalman smoother for a USER vector field `f(t, y)` (any torch function): the reference's
    `sequential_eks_solve` takes any `f` (solver.py:76-96).  The filter relinearises at the PREDICTED mean of every
    step, so the linearisation points have to be found one step after the other: a short forward recursion of tiny
    torch ops on the device (autodiff Jacobian via torch.func.jacfwd, reference observations.py:35-40;
    sequential_filtsmooth/filter.py:9-30, 60-92).  Given those (H_k, c_k), the extended Kalman smoother IS the linear
    filter + smoother pass, which the CUDA kernels compute (`run_pass`).  O(N) launches: the baseline path."""
    from ..parallel_filtsmooth import run_pass

    lin = setup["om"].f._pof_lin
    f, E0, E1 = lin["f"], lin["E0"], lin["E1"]
    d, q = lin["d"], lin["q"]
    D = d * (q + 1)
    dev = setup["_device"]
    F, QL = setup["dtm"].F, setup["dtm"].QL
    x0 = setup["x0"]
    N = len(setup["ts"])
    n = N - 1
    H = torch.empty((n, d, D), dtype=torch.float64, device=dev)
    c = torch.empty((n, d), dtype=torch.float64, device=dev)
    fy = lambda y: f(None, y)
    jac = torch.func.jacfwd(fy)
    zdd = torch.zeros((d, d), dtype=torch.float64, device=dev)
    zDd = torch.zeros((D, d), dtype=torch.float64, device=dev)
    m, L = x0.mean, x0.chol
    for k in range(n):
        mp = F @ m
        Lp = _tria(torch.cat([F @ L, QL], dim=1))
        y = E0 @ mp
        Hk = E1 - jac(y) @ E0
        ck = (E1 @ mp - fy(y)) - Hk @ mp
        H[k], c[k] = Hk, ck
        T = _tria(torch.cat([torch.cat([Hk @ Lp, zdd], dim=1), torch.cat([Lp, zDd], dim=1)], dim=0))
        S, K, L = T[:d, :d], T[d:, :d], T[d:, d:]
        m = mp - K @ torch.linalg.solve_triangular(S, (Hk @ mp + ck).unsqueeze(-1), upper=False).squeeze(-1)
    means = torch.zeros((N, D), dtype=torch.float64, device=dev)
    chols = torch.empty((N, D, D), dtype=torch.float64, device=dev)
    sc = run_pass(x0, setup["_qL"], H, c, means, chols, d=d, q=q, calibrate=False).cpu()
    # the sequential path's `ell` is +sum log-likelihood (filter.py:91), the pass returns the negative sum
    return MVNSqrt(means, chols), -float(sc[nat.S_NLL]), float(sc[nat.S_OBJ]), float(sc[nat.S_SSQ])
