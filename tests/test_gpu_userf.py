"""User-supplied vector fields `f(t, y)` (any torch function, no fused CUDA linearisation): the reference's API takes any
`f` (solver.py:11-96).  The Jacobians come from torch.func autodiff on the device; everything after the linearisation
is the same CUDA pass.  Each path is checked against the built-in problem with the same vector field and the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ivps as oivps  # noqa: E402
from oracle import pof_oracle as O  # noqa: E402


def _user(ivp):
    f = ivp.f
    return lambda t, y: f(t, y)  # a plain function: no `_pof_builtin` attribute -> autodiff linearisation


@pytest.mark.parametrize("name,kw,N,q", [("logistic", {}, 21, 3), ("lotkavolterra", {}, 150, 2),
                                         ("fitzhughnagumo", {}, 120, 3)])
def test_sequential_eks_solve_user_f(native_lib, name, kw, N, q):
    import pof.ivp
    from pof.solver import sequential_eks_solve

    ivp, oivp = getattr(pof.ivp, name)(**kw), getattr(oivps, name)(**kw)
    ts = np.linspace(ivp.t0, ivp.tmax, N)
    g = _user(ivp)
    assert getattr(g, "_pof_builtin", None) is None
    ys, info = sequential_eks_solve(f=g, y0=ivp.y0, ts=ts, order=q)
    oys, oinfo = O.sequential_eks_solve(oivp, ts, q)
    y, yo = ys.mean.cpu().numpy(), oys.mean
    assert (np.abs(y - yo) <= 1e-9 * np.abs(yo).max(axis=0) + 1e-12).all()
    s, so = info["sigma_squared"], oinfo["sigma_squared"]
    C = (ys.chol @ ys.chol.transpose(-1, -2)).cpu().numpy() / s
    Co = oys.chol @ np.swapaxes(oys.chol, -1, -2) / so
    assert np.abs(C - Co).max() <= 1e-7 * np.abs(Co).max()
    assert abs(info["nll"] - oinfo["nll"]) <= 1e-9 * abs(oinfo["nll"]) + 1e-9
    assert abs(s - so) <= 5e-2 * abs(so)


def test_solve_user_f_matches_builtin(native_lib):
    """solve() with a user f (autodiff linearisation + dense-H pass) == solve() with the fused built-in linearisation"""
    import pof.ivp
    from pof.solver import solve

    ivp = pof.ivp.rigid_body()  # (converges from every initialisation: 10 iterations)
    ts = np.linspace(ivp.t0, ivp.tmax, 400)
    a, ia = solve(f=ivp.f, y0=ivp.y0, ts=ts, order=3, init="constant", maxiters=200)
    b, ib = solve(f=_user(ivp), y0=ivp.y0, ts=ts, order=3, init="constant", maxiters=200)
    assert ia["iterations"] == ib["iterations"]
    assert (a.mean - b.mean).abs().max().item() <= 1e-9 * a.mean.abs().max().item()
    # init="coarse" runs a sequential EKS on a coarse grid: needs the user-f sequential path
    c, ic = solve(f=_user(ivp), y0=ivp.y0, ts=ts, order=3, init="coarse", maxiters=200)
    assert (a.mean - c.mean).abs().max().item() <= 1e-6 * a.mean.abs().max().item()


def test_solve_sharded_user_f_single_rank(native_lib):
    """pof.sharded.solve_sharded with a user f (world size 1: the three shard stages with dense (H, c))"""
    import pof.ivp
    from pof.sharded import solve_sharded
    from pof.solver import solve

    ivp = pof.ivp.rigid_body()
    ts = np.linspace(ivp.t0, ivp.tmax, 1500)
    ys, info, rows = solve_sharded(f=_user(ivp), y0=ivp.y0, ts=ts, order=2, init="constant", maxiters=100)
    ref, rinfo = solve(f=ivp.f, y0=ivp.y0, ts=ts, order=2, init="constant", maxiters=100)
    assert info["iterations"] == rinfo["iterations"]
    assert (ys.mean - ref.mean[rows]).abs().max().item() <= 1e-9 * ref.mean.abs().max().item()


def test_solve_batch_equals_separate_solves(native_lib):
    """a batch of independent IVPs advanced in lockstep on separate streams (pof.batch.solve_batch) gives exactly what
    separate `solve` calls give: same iteration counts, same trajectories"""
    import pof.ivp
    from pof.batch import solve_batch
    from pof.solver import solve

    probs = []
    for i, (name, N) in enumerate([("logistic", 200), ("rigid_body", 700), ("vanderpol", 300), ("logistic", 64),
                                   ("rigid_body", 257), ("fitzhughnagumo", 150)]):
        ivp = getattr(pof.ivp, name)() if name != "vanderpol" else pof.ivp.vanderpol(stiffness_constant=1.0)
        y0 = ivp.y0 * (1.0 + 0.01 * i)
        probs.append(dict(f=ivp.f, y0=y0, ts=np.linspace(ivp.t0, ivp.tmax, N)))
    # orders differ per call in the reference's runner; one order per batch here
    res = solve_batch(probs, order=3, init="constant", maxiters=300)
    torch.cuda.synchronize()
    for p, (ys, info) in zip(probs, res):
        ref, rinfo = solve(f=p["f"], y0=p["y0"], ts=p["ts"], order=3, init="constant", maxiters=300)
        assert info["iterations"] == rinfo["iterations"]
        assert torch.equal(ys.mean, ref.mean) and torch.equal(ys.chol, ref.chol)


def test_reference_benchmark_call_shapes(native_lib):
    """The calls of the reference's single-step benchmark (experiments/1_single_step_runtime_gpus/collect_data.py:22-42):
    `linearize_observation_model(om, states[1:])` + `linear_filtsmooth`, and `seq_fs(x0, dtm, om)`, on the CUDA paths."""
    import pof.ivp
    from pof.convenience import get_initial_trajectory, linearize_observation_model, set_up_solver
    from pof.parallel_filtsmooth import linear_filtsmooth
    from pof.sequential_filtsmooth import filtsmooth as seq_fs
    from pof.sequential_filtsmooth.eks import eks_filtsmooth
    from pof.step import linearize_at_previous_states
    from pof.utils import MVNSqrt

    ivp = pof.ivp.fitzhughnagumo()
    N = 64
    setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=np.linspace(0, 5, N), order=3)
    x0, dtm, om = setup["x0"], setup["dtm"], setup["om"]
    states = get_initial_trajectory(setup, method="constant")
    dom = linearize_observation_model(om, MVNSqrt(states.mean[1:], states.chol[1:]))
    dom2 = linearize_at_previous_states(om, states)
    assert dom.H.shape == (N - 1, 2, 8) and torch.equal(dom.H, dom2.H) and torch.equal(dom.b, dom2.b)
    out, nll, obj, ssq = linear_filtsmooth(x0, dtm, dom)
    assert out.mean.shape == (N, 8) and bool(torch.isfinite(out.mean).all())
    s1, ell1, obj1, ssq1 = seq_fs(x0, dtm, om, n=N - 1)
    s2, ell2, obj2, ssq2 = eks_filtsmooth(setup)
    assert torch.equal(s1.mean, s2.mean) and (ell1, obj1, ssq1) == (ell2, obj2, ssq2)
    with pytest.raises(ValueError):
        seq_fs(x0, dtm, om)  # one (D,D) copy of the model: the number of steps has to be given
