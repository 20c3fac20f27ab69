"""The IEKS loop on the device (reference pof/solver.py:36-57, `jax.lax.while_loop`): `solve` runs it as ONE CUDA graph
with a WHILE conditional node (`pof_ieks_loop_create`) or, as a fallback, as bursts of captured single iterations
(`pof_ieks_loop_step`); both must stop at exactly the iteration a host-side loop that checks the reference's rule after
every iteration stops at, with identical results."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _host_loop(ivp, ts, q, maxiters):
    """the loop written out on the host with the single-iteration entry point and the reference's rule"""
    from pof import _native as nat
    from pof.convenience import get_initial_trajectory, set_up_solver
    from pof.convergence_criteria import crit_scalars
    from pof.parallel_filtsmooth import run_iteration

    setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=ts, order=q)
    lin = setup["om"].f._pof_lin
    means = get_initial_trajectory(setup, method="constant", means_only=True).mean.contiguous()
    chols = torch.empty((means.shape[0], means.shape[1], means.shape[1]), dtype=means.dtype, device=means.device)
    scalars = torch.zeros(nat.NSCALARS, dtype=means.dtype, device=means.device)
    k, nll, obj, nll_old, obj_old, bad = 0, 0.0, 0.0, 0.0, 0.0, 1.0
    while True:
        if k >= 1 and (crit_scalars(obj, obj_old, nll, nll_old, bad) or not (k <= maxiters)):
            break
        nll_old, obj_old = nll, obj
        run_iteration(setup["x0"], setup["_qL"], lin, means, chols, calibrate=True, scalars=scalars)
        sc = scalars.cpu()
        nll, obj, bad = float(sc[nat.S_NLL]), float(sc[nat.S_OBJ]), float(sc[nat.S_NOT_CLOSE])
        k += 1
    return k, nll, obj, setup["_scale0"] * means[:, 0]


@pytest.mark.parametrize("name,N,q,maxiters", [("logistic", 64, 2, 10_000), ("fitzhughnagumo", 512, 3, 10_000),
                                               ("fitzhughnagumo", 4096, 3, 10_000), ("rigid_body", 300, 4, 10_000),
                                               ("fitzhughnagumo", 512, 3, 3), ("fitzhughnagumo", 512, 3, 0)])
def test_device_loop_stops_where_the_host_loop_stops(native_lib, monkeypatch, name, N, q, maxiters):
    import pof.ivp
    from pof import _native as nat
    from pof.solver import solve

    ivp = getattr(pof.ivp, name)()
    ts = np.linspace(ivp.t0, ivp.tmax, N)
    k_ref, nll_ref, obj_ref, y0_ref = _host_loop(ivp, ts, q, maxiters)
    if maxiters < 10:
        assert k_ref == maxiters + 1  # quirk Q4: the reference runs maxiters + 1 iterations
    res = {}
    for graph in (True, False):
        monkeypatch.setattr(nat, "USE_LOOP_GRAPH", graph)
        ys, info = solve(f=ivp.f, y0=ivp.y0, ts=ts, order=q, init="constant", maxiters=maxiters)
        assert info["iterations"] == k_ref
        assert info["nll"] == nll_ref and info["obj"] == obj_ref  # same kernels, same order: bit-identical
        res[graph] = ys
    assert torch.equal(res[True].mean, res[False].mean)
    assert torch.equal(res[True].chol, res[False].chol)
    assert torch.allclose(res[True].mean[:, 0], y0_ref, rtol=1e-15, atol=0)  # E0 projection: a constant scaling


def test_loop_graph_entry_point_directly(native_lib):
    """the C entry points: create after one eager step, launch, read the loop state once; a second launch with the
    stop flag still set is a no-op"""
    import pof.ivp
    from pof import _native as nat
    from pof.convenience import get_initial_trajectory, set_up_solver
    from pof.parallel_filtsmooth import GraphedIteration

    ivp = pof.ivp.fitzhughnagumo()
    ts = np.linspace(0, 20, 1000)
    setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=ts, order=3)
    lin = setup["om"].f._pof_lin
    means = get_initial_trajectory(setup, method="constant", means_only=True).mean.contiguous()
    chols = torch.empty((1000, 8, 8), dtype=torch.float64, device=means.device)
    scalars = torch.zeros(nat.NSCALARS, dtype=torch.float64, device=means.device)
    ls = torch.zeros(8, dtype=torch.float64, device=means.device)
    it = GraphedIteration(setup["x0"], setup["_qL"], lin, means, chols, scalars, loop_state=ls, maxiters=10_000)
    it()
    assert float(ls[1]) == 1.0 and float(ls[0]) == 0.0
    assert it.capture_loop()
    it.launch_loop()
    torch.cuda.synchronize()
    k = float(ls[1])
    assert float(ls[0]) == 1.0 and 3 < k < 200
    snapshot = means.clone()
    it.launch_loop()  # stop flag set: nothing may change
    torch.cuda.synchronize()
    assert float(ls[1]) == k and torch.equal(means, snapshot)
    ls.zero_()  # re-armed: continues from the converged trajectory and stops after very few iterations
    it.launch_loop()
    torch.cuda.synchronize()
    assert float(ls[0]) == 1.0 and 1 <= float(ls[1]) <= 3
    # (the loop had stopped on the objective rule, rtol 1e-6: the extra iterations still move the means a little)
    assert bool(((means - snapshot).abs().amax(0) <= 1e-5 * snapshot.abs().amax(0)).all())


def test_loop_graph_fp32(native_lib, monkeypatch):
    import pof.ivp
    from pof import _native as nat
    from pof.solver import solve

    ivp = pof.ivp.logistic()
    ts = np.linspace(ivp.t0, ivp.tmax, 256)
    out = {}
    for graph in (True, False):
        monkeypatch.setattr(nat, "USE_LOOP_GRAPH", graph)
        ys, info = solve(f=ivp.f, y0=ivp.y0, ts=ts, order=2, init="constant", maxiters=200, dtype=torch.float32)
        out[graph] = (ys.mean, info["iterations"])
    assert out[True][1] == out[False][1]
    assert torch.equal(out[True][0], out[False][0])
