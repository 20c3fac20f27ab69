"""Generator API of the IEKS and its regularised variant (reference pof/iterators.py:20-112).

`ieks_iterator` yields the iterates of the plain loop; `qpm_ieks_iterator` is the reference's quadratic-penalty
method: the ODE residual is observed with noise R = (reg / n) I that is driven geometrically from `reg_start` to
`reg_final` and then to zero, every stage iterated to a tolerance `tau` that tightens along -- the robustness variant
for problems on which the plain Gauss-Newton iteration diverges.  The noisy passes (cholR != 0) run on the
large-state CUDA kernels (`pof_linear_filtsmooth_general_f64`); nothing is computed on the host.

Not provided: `lm_ieks_iterator` needs observations of dimension d + D (reference observations.py:66-83,
`linearize_regularized`), which the kernels -- whose observation dimension is the ODE dimension d -- do not take;
`linearize_regularized` itself is available in pof.observations and `lm_ieks_iterator` raises NotImplementedError.
"""
import torch

from . import convergence_criteria
from .convenience import get_initial_trajectory, set_up_solver
from .observations import AffineModel
from .parallel_filtsmooth import linear_filtsmooth
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


def lm_ieks_iterator(dtm, om, x0, init_traj, reg=1e0, nu=10.0):
    """reference iterators.py:115-138 (Levenberg-Marquardt: observations of dimension d + D)"""
    raise NotImplementedError(
        "lm_ieks_iterator: the regularised observation model has dimension d + D (pof.observations."
        "linearize_regularized); the CUDA kernels take observations of the ODE dimension d only -- use "
        "qpm_ieks_iterator for a regularised iteration")


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
