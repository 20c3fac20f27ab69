"""Optional fp32 mode (BASELINE north_star: "an optional fp32 mode reported separately"): the register-resident kernels
compiled with the scalar type float, against the fp64 oracle.  fp32 is only usable at small N (the disagreement of
valid association orders already grows like N^3 in fp64): gated at 1e-4 on the outputs for N <= 256
from the constant trajectory; scripts/fp32_accuracy.py maps out where it breaks (profiles/r02_fp32_accuracy.md)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ivps as oivps  # noqa: E402
from oracle import pof_oracle as O  # noqa: E402


@pytest.mark.parametrize("name,kw,N,q", [("fitzhughnagumo", {}, 100, 3), ("logistic", {}, 64, 2),
                                         ("lotkavolterra", {}, 256, 2), ("rigid_body", {}, 200, 3)])
def test_fp32_iteration_matches_oracle(native_lib, name, kw, N, q):
    import pof.ivp
    from pof import _native as nat
    from pof.convenience import get_initial_trajectory, set_up_solver
    from pof.parallel_filtsmooth import run_iteration
    from pof.utils import MVNSqrt

    ivp, oivp = getattr(pof.ivp, name)(**kw), getattr(oivps, name)(**kw)
    ts = np.linspace(ivp.t0, ivp.tmax, N)
    setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=ts, order=q)
    f32 = torch.float32
    means = get_initial_trajectory(setup, method="constant", means_only=True).mean.to(f32).contiguous()
    D = means.shape[1]
    chols = torch.empty((N, D, D), dtype=f32, device=means.device)
    x0 = MVNSqrt(setup["x0"].mean.to(f32), setup["x0"].chol.to(f32))
    sc = run_iteration(x0, setup["_qL"], setup["om"].f._pof_lin, means, chols, calibrate=False)
    torch.cuda.synchronize()
    assert sc.dtype == f32 and means.dtype == f32
    sc = sc.cpu().numpy().astype(np.float64)

    osetup = O.set_up_solver(oivp, ts, q)
    oout, onll, oobj, ossq, ossqp = O.ieks_step(osetup, O.get_initial_trajectory(osetup), calibrate=False)
    E0 = osetup["E0"]
    y, yo = means.cpu().numpy().astype(np.float64) @ E0.T, oout.mean @ E0.T
    err = np.abs(y - yo).max(axis=0) / np.abs(yo).max(axis=0)
    assert (err <= 1e-4).all(), err
    Lg = chols.cpu().numpy().astype(np.float64)
    Pg = E0 @ (Lg @ np.swapaxes(Lg, -1, -2)) @ E0.T
    Po = E0 @ (oout.chol @ np.swapaxes(oout.chol, -1, -2)) @ E0.T
    assert np.abs(Pg - Po).max() <= 1e-3 * np.abs(Po).max()
    assert abs(sc[nat.S_NLL] - onll) <= 1e-3 * abs(onll) + 1e-2
    assert abs(sc[nat.S_SSQ_PROPER] - ossqp) <= 1e-2 * abs(ossqp)
    assert np.abs(np.triu(Lg, 1)).max() == 0.0


def test_fp32_solve_matches_fp64(native_lib):
    """solve(dtype=float32) on the reference's own test problem (logistic, dt = 0.5, tests/test_solver.py): float32
    outputs within 1e-4 of the fp64 solution; it stops on the objective rule (the 1e-8 mean rule is below fp32 eps)"""
    import pof.ivp
    from pof.solver import solve

    ivp = pof.ivp.logistic()
    ts = np.arange(0, ivp.tmax + 0.5, 0.5)
    a, ia = solve(f=ivp.f, y0=ivp.y0, ts=ts, order=3, init="constant")
    b, ib = solve(f=ivp.f, y0=ivp.y0, ts=ts, order=3, init="constant", dtype=torch.float32, maxiters=50)
    assert b.mean.dtype == torch.float32 and b.chol.shape == a.chol.shape
    assert (a.mean - b.mean.double()).abs().max().item() <= 1e-4 * a.mean.abs().max().item()
    assert ib["iterations"] <= 51


def test_fp32_rejects_what_it_does_not_cover(native_lib):
    import pof.ivp
    from pof.solver import solve

    ivp = pof.ivp.lorenz96(tmax=0.2, d=8)
    with pytest.raises(Exception):
        solve(f=ivp.f, y0=ivp.y0, ts=np.linspace(0, 0.2, 50), order=3, init="constant", dtype=torch.float32)


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("N", [256, 5000, 70_000])
def test_pass_stays_inside_its_workspace(native_lib, dtype, N):
    """`pof_workspace_bytes[_f32]` must cover everything a pass touches (the dataflow flag area is 32-bit words in a
    buffer laid out in units of the scalar type): canary bytes behind the workspace survive an iteration"""
    import pof.ivp
    from pof import _native as nat
    from pof.convenience import get_initial_trajectory, set_up_solver
    from pof.parallel_filtsmooth import run_iteration
    from pof.utils import MVNSqrt

    ivp = pof.ivp.fitzhughnagumo()
    setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=np.linspace(0, 10, N), order=3)
    lin = setup["om"].f._pof_lin
    x0 = MVNSqrt(setup["x0"].mean.to(dtype), setup["x0"].chol.to(dtype))
    means = get_initial_trajectory(setup, method="constant", means_only=True).mean.to(dtype).contiguous()
    chols = torch.empty((N, 8, 8), dtype=dtype, device=means.device)
    L = nat.default_chunk_len(N, 2, 3, means.device.index)
    ws = nat.Workspace(N, 2, 3, L, means.device, dtype)
    tail = 1 << 16
    big = torch.full((ws.nbytes + tail,), 0xAB, dtype=torch.uint8, device=means.device)
    ws.buf = big  # same size handed to the library, canary behind it
    for _ in range(2):
        run_iteration(x0, setup["_qL"], lin, means, chols, calibrate=True, chunk_len=L, ws=ws)
    torch.cuda.synchronize()
    assert bool((big[ws.nbytes:] == 0xAB).all())
    assert bool(torch.isfinite(means).all())
