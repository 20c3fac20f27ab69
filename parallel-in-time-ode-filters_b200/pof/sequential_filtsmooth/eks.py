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
        raise NotImplementedError(
            "sequential_eks_solve is implemented for the built-in pof.ivp vector fields only (the per-step "
            "relinearisation at the predicted mean runs inside the kernel)")
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
