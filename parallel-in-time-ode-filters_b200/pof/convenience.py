"""Problem set-up (reference pof/convenience.py:13-45, 76-92)."""
import numpy as np
import torch

from . import initialization as init
from .observations import NonlinearModel
from .transitions import (IWP, TransitionModel, discretize_transitions, nordsieck_preconditioner, nordsieck_scalings,
                          preconditioned_discretize, preconditioned_discretize_1d, projection_matrix)
from .utils import MVNSqrt


def linearize_observation_model(observation_model, trajectory):
    """reference convenience.py:9-10: the EK1 linearisation (observations.py:35-40) at EVERY state of `trajectory`
    (MVNSqrt with means (n, D)) -> AffineModel(H (n,d,D), b (n,d), cholR (n,d,d) = 0).  Models made by `set_up_solver`
    for a built-in `pof.ivp` problem use the fused CUDA kernel, any other callable torch autodiff on the means' device."""
    from .observations import AffineModel
    from .step import linearize_at_previous_states

    means = trajectory.mean
    if getattr(observation_model.f, "_pof_lin", None) is not None:
        # `linearize_at_previous_states` linearises at rows 1..n of a trajectory: put a dummy row in front
        padded = MVNSqrt(torch.cat([means[:1], means]), None)
        return linearize_at_previous_states(observation_model, padded)
    f = observation_model.f
    H = torch.func.vmap(torch.func.jacfwd(f))(means)
    res = torch.func.vmap(f)(means)
    b = res - torch.einsum("nij,nj->ni", H, means)
    n, d = b.shape
    return AffineModel(H, b, torch.zeros((n, d, d), dtype=means.dtype, device=means.device))


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("pof_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def set_up_solver(*, f, y0, ts, order):
    """reference convenience.py:13-45.  `dtm` holds ONE (D,D) copy of F and QL (the reference replicates the
    identical matrices n times); `om.f` carries what the fused linearisation kernel needs."""
    dev = _device()
    ts_h = np.asarray(torch.as_tensor(ts, dtype=torch.float64).detach().cpu())
    y0_h = torch.as_tensor(y0, dtype=torch.float64).detach().cpu()
    dt = float(ts_h[1] - ts_h[0])  # uniform grid assumed, as in the reference (convenience.py:14-16)
    d = int(y0_h.shape[0])

    iwp = IWP(num_derivatives=order, wiener_process_dimension=d)
    F, QL = preconditioned_discretize(iwp)
    _, qL = preconditioned_discretize_1d(iwp)
    P, PI = nordsieck_preconditioner(iwp, dt)
    sv, _ = nordsieck_scalings(iwp, dt)
    E0 = projection_matrix(iwp, 0) @ P
    E1 = projection_matrix(iwp, 1) @ P

    tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    E0_t, E1_t = tt(E0), tt(E1)

    def om_f(x):
        return E1_t.to(x.device) @ x - f(None, E0_t.to(x.device) @ x)

    om_f._pof_lin = dict(
        builtin=getattr(f, "_pof_builtin", None), f=f, scale0=float(sv[0]), scale1=float(sv[1]) if order >= 1 else 0.0,
        d=d, q=order, E0=E0_t, E1=E1_t,
    )
    om = NonlinearModel(om_f)

    x0_raw = init.taylor_mode_init(f, y0_h, order)
    PI_t = tt(PI)
    x0 = MVNSqrt(PI_t @ x0_raw.mean.to(dev), PI_t @ x0_raw.chol.to(dev))

    return {
        "f": f, "y0": y0_h, "ts": ts_h, "dtm": TransitionModel(tt(F), tt(QL)), "om": om, "x0": x0,
        "E0": E0_t, "P": tt(P), "PI": PI_t, "order": order, "iwp": iwp,
        "_qL": np.ascontiguousarray(qL), "_scale0": float(sv[0]), "_device": dev, "_x0_raw": x0_raw,
    }


def set_up_solver_no_precond(*, f, y0, ts, order):
    """reference convenience.py:48-73: non-preconditioned coordinates, one dense transition model per step -- valid on
    NON-uniform grids.  `dtm` holds (n,D,D) stacks; `linear_filtsmooth` / `ieks_step` route them to the general pass."""
    dev = _device()
    ts_h = np.asarray(torch.as_tensor(ts, dtype=torch.float64).detach().cpu())
    y0_h = torch.as_tensor(y0, dtype=torch.float64).detach().cpu()
    d = int(y0_h.shape[0])
    iwp = IWP(num_derivatives=order, wiener_process_dimension=d)
    dtm = discretize_transitions(iwp, times=ts_h, device=dev)
    tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    E0_t, E1_t = tt(projection_matrix(iwp, 0)), tt(projection_matrix(iwp, 1))

    def om_f(x):
        return E1_t.to(x.device) @ x - f(None, E0_t.to(x.device) @ x)

    om_f._pof_lin = dict(builtin=getattr(f, "_pof_builtin", None), f=f, scale0=1.0, scale1=1.0, d=d, q=order,
                         E0=E0_t, E1=E1_t)
    x0 = init.taylor_mode_init(f, y0_h, order)
    x0 = MVNSqrt(x0.mean.to(dev), x0.chol.to(dev))
    eye = torch.eye(d * (order + 1), dtype=torch.float64, device=dev)
    _, qL = preconditioned_discretize_1d(iwp)
    return {
        "f": f, "y0": y0_h, "ts": ts_h, "dtm": dtm, "om": NonlinearModel(om_f), "x0": x0, "E0": E0_t, "P": eye,
        "PI": eye, "order": order, "iwp": iwp, "_qL": np.ascontiguousarray(qL), "_scale0": 1.0, "_device": dev,
    }


def get_initial_trajectory(setup, method="prior", *, means_only=False):
    """reference convenience.py:76-92.  Everything is built on the device; `means_only` (used by `solve`: only the
    means feed the first linearisation) skips the (N,D,D) Cholesky factors (`chol` is then None)."""
    f, y0, order, ts = setup["f"], setup["y0"], setup["order"], setup["ts"]
    PI, dev = setup["PI"], setup["_device"]
    if method == "coarse":
        st = init.coarse_ekf_init(y0=y0, order=order, ts=ts, f=f, N=100)
        chol = None if means_only else torch.einsum("ij,njk->nik", PI, st.chol.to(dev))
        return MVNSqrt((st.mean.to(dev) @ PI.T).contiguous(), chol)
    elif method == "constant":
        N = len(ts)
        row = init.constant_init(y0=y0, order=order, ts=ts[:1], f=f).mean.to(dev)  # (1, D): every row is the same
        means = (row @ PI.T).expand(N, -1).contiguous()
        D = means.shape[1]
        chol = None if means_only else torch.zeros((N, D, D), dtype=torch.float64, device=dev)  # PI @ 0 = 0
        return MVNSqrt(means, chol)
    elif method == "prior":
        return init.prior_init(f=f, y0=y0, order=order, ts=ts, device=dev, means_only=means_only,
                               x0=setup.get("_x0_raw"))
    else:
        raise Exception(f"method={method} not found")
