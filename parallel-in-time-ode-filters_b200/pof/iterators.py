"""Generator API of the IEKS and its regularised variant (reference pof/iterators.py:20-112).

`ieks_iterator` yields the iterates of the plain loop; `qpm_ieks_iterator` is the reference's quadratic-penalty
method: the ODE residual is observed with noise R = (reg / n) I that is driven geometrically from `reg_start` to
`reg_final` and then to zero, every stage iterated to a tolerance `tau` that tightens along -- the robustness variant
for problems on which the plain Gauss-Newton iteration diverges.  The noisy passes (cholR != 0) run on the
large-state CUDA kernels (`pof_linear_filtsmooth_general_f64`); nothing is computed on the host.

`lm_ieks_iterator` (upstream: uncalled, work in progress) stacks the EK1 model with a pseudo-observation of the whole
state (reference observations.py:65-83): observation dimension d + D, which the kernels -- whose observation dimension
is the ODE dimension d -- do not take.  Its passes run through torch's batched library calls on the device
(`pof.parallel_filtsmooth.library_pass`): functional, not a performance path.
"""
import torch

from . import convergence_criteria
from .convenience import get_initial_trajectory, set_up_solver
from .observations import AffineModel
from .parallel_filtsmooth import linear_filtsmooth
from .parallel_filtsmooth.library_pass import linear_filtsmooth_library
from .step import ieks_step, linearize_at_previous_states


def ieks_iterator(*, f, y0, ts, order, init="prior"):
    """reference iterators.py:20-24"""
    setup = set_up_solver(f=f, y0=y0, ts=ts, order=order)
    states = get_initial_trajectory(setup, method=init)
    return _ieks_iterator(setup["dtm"], setup["om"], setup["x0"], states), setup


def _ieks_iterator(dtm, om, x0, init_traj):
    """reference iterators.py:27-41"""
    states, nll, obj, ssq = ieks_step(om=om, dtm=dtm, x0=x0, states=init_traj)
    yield states, nll, obj, ssq
    while True:
        nll_old, obj_old, states_old = nll, obj, states
        states, nll, obj, ssq = ieks_step(om=om, dtm=dtm, x0=x0, states=states_old)
        yield states, nll, obj, ssq
        if convergence_criteria.crit(obj, obj_old, nll, nll_old, states, states_old):
            break


def qpm_ieks_iterator(*, f, y0, ts, order, init="prior", **kwargs):
    """reference iterators.py:44-50"""
    setup = set_up_solver(f=f, y0=y0, ts=ts, order=order)
    states = get_initial_trajectory(setup, method=init)
    return _qpm_ieks_iterator(setup["dtm"], setup["om"], setup["x0"], states, **kwargs), setup


def _qpm_ieks_iterator(dtm, om, x0, init_traj, reg_start=1e20, reg_final=1e-20, steps=40, tau_start=None,
                       tau_final=None):
    """reference iterators.py:53-112.  Yields (states, nll, obj, reg)."""
    dom = linearize_at_previous_states(om, init_traj)
    n, d = dom.b.shape  # the reference's N = dtm.F.shape[0] = number of transitions
    reg_fact = (reg_final / reg_start) ** (1 / steps)
    if tau_start is None:
        tau_start, tau_final = 1e5, 1e-5
    tau_fact = (tau_final / tau_start) ** (1 / steps)
    reg, tau = reg_start, tau_start
    eye = torch.eye(d, dtype=torch.float64, device=dom.H.device).expand(n, d, d)

    def noisy(dom, reg):
        return AffineModel(dom.H, dom.b, (reg / n) * eye)

    states, nll, obj, ssq = linear_filtsmooth(x0, dtm, noisy(dom, reg))
    yield states, nll, obj, reg
    while True:
        nll_old, obj_old, states_old = nll, obj, states
        dom = linearize_at_previous_states(om, states)
        states, nll, obj, ssq = linear_filtsmooth(x0, dtm, noisy(dom, reg))
        yield states, nll, obj, reg
        if convergence_criteria.crit(obj, obj_old, nll, nll_old, states, states_old, rtol=tau, atol=tau):
            reg *= reg_fact
            tau *= tau_fact
            if reg == 0:
                break
            elif reg < reg_final:
                reg = 0.0
                tau = min(tau_final, 1e-5)
        if bool(torch.isnan(torch.as_tensor(nll))) or bool(torch.isnan(torch.as_tensor(obj))):
            break


def stack_regularized(dom, means, reg):
    """Batched `linearize_regularized` (reference observations.py:65-83) from the EK1 model `dom` linearised at `means`
    (n, D): the ODE observation stacked with the pseudo-observation x ~ N(m, I / reg); observation dimension d + D.
    Kept with the reference's (work-in-progress) offsets `full_b = [f(m), -m]` -- NOT the EK1 offset f(m) - H m."""
    H, b = dom.H, dom.b
    n, d, D = H.shape
    res = b + torch.einsum("nij,nj->ni", H, means)  # EK1: b = f(m) - H m
    eye = torch.eye(D, dtype=H.dtype, device=H.device).expand(n, D, D)
    cholR = torch.zeros((n, d + D, d + D), dtype=H.dtype, device=H.device)
    cholR[:, d:, d:] = eye / reg ** 0.5
    return AffineModel(torch.cat([H, eye], dim=1), torch.cat([res, -means], dim=1), cholR)


def lm_ieks_iterator(dtm, om, x0, init_traj, reg=1e0, nu=10.0):
    """reference iterators.py:109-133 (Levenberg-Marquardt-style regularisation; upstream the accept / reject step that
    would use `nu` is commented out, so `reg` stays constant).  Yields (states, nll, obj, reg); stops when nll AND obj
    are `isclose` (rtol 1e-5, atol 1e-8) to the previous iterate's, or on NaN.  Every trajectory is linearised at its
    states 1..n (the upstream first call vmaps over all of `init_traj`, which only fits a trajectory of n states)."""

    def one(states):
        means = states.mean[1:]
        dom = stack_regularized(linearize_at_previous_states(om, states), means, reg)
        out, nll, obj, _ = linear_filtsmooth_library(x0, dtm, dom)
        return out, nll, obj

    isclose = lambda a, b: bool(torch.isclose(torch.as_tensor(a), torch.as_tensor(b), rtol=1e-5, atol=1e-8))
    out, nll, obj = one(init_traj)
    yield out, nll, obj, reg
    while True:
        nll_old, obj_old = nll, obj
        out, nll, obj = one(out)
        yield out, nll, obj, reg
        if (isclose(nll_old, nll) and isclose(obj_old, obj)) or bool(torch.isnan(torch.as_tensor(nll))) or bool(
                torch.isnan(torch.as_tensor(obj))):
            break


def admm_ieks_iterator(*, f, y0, ts, order, init="prior"):
    """reference iterators.py:141-157: as committed upstream this is the plain IEKS loop (`rho` is unused) with the
    stopping rule evaluated on the scalars only"""
    setup = set_up_solver(f=f, y0=y0, ts=ts, order=order)
    states = get_initial_trajectory(setup, method=init)

    def gen():
        st = states
        st, nll, obj, _ = ieks_step(om=setup["om"], dtm=setup["dtm"], x0=setup["x0"], states=st, calibrate=False)
        yield st, nll, obj
        while True:
            nll_old, obj_old = float(nll), float(obj)
            st, nll, obj, _ = ieks_step(om=setup["om"], dtm=setup["dtm"], x0=setup["x0"], states=st, calibrate=False)
            yield st, nll, obj
            if convergence_criteria.crit_scalars(float(obj), obj_old, float(nll), nll_old, 1):
                break

    return gen(), setup
