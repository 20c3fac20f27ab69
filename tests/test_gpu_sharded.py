"""Time-sharded pass on real GPUs (NCCL): sharded == single-GPU pass == oracle.  Needs >= 2 GPUs (skipped otherwise)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    """rendezvous file for the file:// init method (a probed-free TCP port was taken again before rank 0 listened on it
    once, which cost a 300 s hang on the GPU box)"""
    import tempfile

    fd, path = tempfile.mkstemp(prefix="pof_rdzv_")
    os.close(fd)
    os.unlink(path)
    return path


def _worker(rank, world, port, N, q, ret, exchange="nccl"):
    import torch.distributed as dist

    sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{port}", rank=rank, world_size=world, device_id=dev)
    try:
        import pof.ivp
        from pof.convenience import get_initial_trajectory, set_up_solver
        from pof.parallel_filtsmooth import linear_filtsmooth
        from pof.sharded import ShardedPass, shard_bounds
        from pof.step import linearize_at_previous_states

        ivp = pof.ivp.fitzhughnagumo()
        ts = np.linspace(0, 100, N)
        setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=ts, order=q)
        st = get_initial_trajectory(setup, method="constant")
        dom = linearize_at_previous_states(setup["om"], st)
        ref, nll, obj, ssq = linear_filtsmooth(setup["x0"], setup["dtm"], dom)
        d, D = 2, 2 * (q + 1)
        k_lo, k_hi = shard_bounds(N - 1, rank, world)
        r0 = 0 if rank == 0 else k_lo + 1
        try:
            sp = ShardedPass(N, d, q, setup["_qL"], rank=rank, world=world, device=dev, exchange=exchange)
        except Exception as e:  # peer memory (CUDA IPC) not available in this environment
            ret[rank] = ("skip", repr(e))
            return
        assert sp.exchange == exchange
        means = st.mean[r0:k_hi + 1].contiguous().clone()
        chols = torch.zeros((sp.rows, D, D), dtype=torch.float64, device=dev)
        res = sp.run(setup["x0"].mean, setup["x0"].chol, dom.H[k_lo:k_hi].contiguous(), dom.b[k_lo:k_hi].contiguous(),
                     means, chols, calibrate=False)
        torch.cuda.synchronize()
        em = float((means - ref.mean[r0:k_hi + 1]).abs().max() / ref.mean.abs().max())
        cov = lambda L: L @ L.transpose(-1, -2)
        ec = float((cov(chols) - cov(ref.chol[r0:k_hi + 1])).abs().max() / cov(ref.chol).abs().max())
        ok = (em < 1e-9 and ec < 1e-9 and abs(float(res["nll"]) - float(nll)) <= 1e-9 * abs(float(nll))
              and abs(float(res["obj"]) - float(obj)) <= 1e-9 * abs(float(obj))
              and abs(float(res["ssq"]) - float(ssq)) <= 1e-6 * abs(float(ssq)))
        if sp.p2p is not None:
            ok = ok and sp.p2p.status() == 0
            # a second pass through the same areas (epochs, double-buffered slots)
            means2 = st.mean[r0:k_hi + 1].contiguous().clone()
            res2 = sp.run(setup["x0"].mean, setup["x0"].chol, dom.H[k_lo:k_hi].contiguous(),
                          dom.b[k_lo:k_hi].contiguous(), means2, chols, calibrate=False)
            torch.cuda.synchronize()
            ok = ok and torch.equal(means, means2) and float(res2["nll"]) == float(res["nll"]) and sp.p2p.status() == 0
        ret[rank] = (bool(ok), em, ec, float(res["nll"]), float(nll))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["nccl", "p2p"])
@pytest.mark.parametrize("N,q", [(4097, 3), (100000, 3)])
def test_sharded_nccl_matches_single_gpu(native_lib, N, q, exchange):
    """exchange = nccl: all-gathers + fused fold kernels; p2p: peer-memory exchange kernels (no collective in the pass)"""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q, ret, exchange)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(180) for p in procs]
    for p in procs:
        if p.is_alive():
            p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    if any(ret[r][0] == "skip" for r in range(world)):
        pytest.skip("peer memory not available: " + str(ret[0]))
    for r in range(world):
        assert ret[r][0] is True, ret[r]


def _solve_worker(rank, world, port, N, q, ret):
    import torch.distributed as dist

    sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{port}", rank=rank, world_size=world, device_id=dev)
    try:
        import pof.ivp
        from pof.sharded import solve_sharded
        from pof.solver import solve

        # rigid body: 10 iterations at every N (Lotka-Volterra at N ~ 2e4 sits at the edge of the IEKS's divergence:
        # there the iteration count depends on the association order of the scan, cf. profiles/r01_work_precision_*)
        ivp = pof.ivp.rigid_body()
        ts = np.linspace(ivp.t0, ivp.tmax, N)
        ys, info, rows = solve_sharded(f=ivp.f, y0=ivp.y0, ts=ts, order=q, init="constant", maxiters=1000)
        ref, rinfo = solve(f=ivp.f, y0=ivp.y0, ts=ts, order=q, init="constant", maxiters=1000)
        torch.cuda.synchronize()
        em = float((ys.mean - ref.mean[rows]).abs().max() / ref.mean.abs().max())
        cov = lambda L: L @ L.transpose(-1, -2)
        ec = float((cov(ys.chol) - cov(ref.chol[rows])).abs().max() / cov(ref.chol).abs().max())
        ok = abs(info["iterations"] - rinfo["iterations"]) <= 1 and em < 1e-7 and ec < 1e-6
        ret[rank] = (bool(ok), info["iterations"], rinfo["iterations"], em, ec)
        torch.cuda.synchronize()
        # leave without tearing the process group down: NCCL kernels captured in CUDA graphs made
        # destroy_process_group hang once (bench.py leaves the same way)
        os._exit(0)
    finally:
        pass


def test_solve_sharded_nccl_matches_single_gpu(native_lib):
    """the sharded IEKS loop (graph-replayed iterations with NCCL all-gathers) against pof.solver.solve"""
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_solve_worker, args=(r, world, port, 20000, 3, ret)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(180) for p in procs]
    for p in procs:
        if p.is_alive():
            p.kill()
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for r in range(world):
        assert ret[r][0], ret[r]


@pytest.mark.parametrize("world", [2, 3, 5, 8])
@pytest.mark.parametrize("name,N,q,compact", [("fitzhughnagumo", 3001, 3, True), ("rigid_body", 1200, 2, False),
                                              ("logistic", 257, 1, True)])
def test_virtual_ranks_on_one_gpu_match_oracle(native_lib, world, name, N, q, compact):
    """The whole time-sharded pass -- three stages per rank, the fused exchange kernels (carry fold + scalar
    bookkeeping) -- with `world` VIRTUAL ranks driven in lockstep on ONE GPU: the all-gathers are concatenations.
    Everything but NCCL itself is the multi-GPU code path, so a single-GPU lease still checks it against the oracle
    (outputs 1e-9, covariances 1e-7, nll / obj 1e-9, sign-invariant sigma^2 1e-8)."""
    import pof.ivp
    from pof import _native as nat
    from pof.convenience import get_initial_trajectory, set_up_solver
    from pof.sharded import ShardedPass
    from pof.step import linearize_at_previous_states

    from oracle import ivps as oivps
    from oracle import pof_oracle as O

    ivp, oivp = getattr(pof.ivp, name)(), getattr(oivps, name)()
    ts = np.linspace(ivp.t0, ivp.tmax, N)
    setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=ts, order=q)
    st = get_initial_trajectory(setup, method="constant")
    lin = setup["om"].f._pof_lin
    d = lin["d"]
    D = d * (q + 1)
    dev = st.mean.device
    dom = linearize_at_previous_states(setup["om"], st)
    sps, states = [], []
    for r in range(world):
        sp = ShardedPass(N, d, q, setup["_qL"], rank=r, world=world, device=dev)
        assert sp.backend.fused_exchange
        r0 = 0 if r == 0 else sp.k_lo + 1
        means = st.mean[r0:sp.k_hi + 1].contiguous().clone()
        chols = torch.zeros((sp.rows, D, D), dtype=torch.float64, device=dev)
        if compact:
            sp.backend.set_compact(lin["scale0"], lin["scale1"])
            Jc = torch.empty((sp.n_loc, d * d + d), dtype=torch.float64, device=dev)
            ivp_id, params = lin["builtin"]
            ph, pp = nat.host_doubles(list(params) + [0.0])
            nat.check(nat.LIB.pof_linearize_ivp_compact_f64(nat.stream_ptr(), ivp_id, pp, len(params), sp.n_loc, d, q,
                                                            lin["scale0"], nat.ptr(st.mean[sp.k_lo + 1:sp.k_hi + 1]
                                                                                   .contiguous()), nat.ptr(Jc)), "lin")
            H, c = Jc, None
        else:
            H, c = dom.H[sp.k_lo:sp.k_hi].contiguous(), dom.b[sp.k_lo:sp.k_hi].contiguous()
        sps.append(sp)
        states.append((setup["x0"].mean, setup["x0"].chol, H, c, means, chols, False, None, None))
    for sp, s in zip(sps, states):
        sp.phase_a(s)
    gf = torch.cat([sp.carry_f for sp in sps])
    for sp, s in zip(sps, states):
        sp.gather_f.copy_(gf)
        sp.phase_b(s)
    gb = torch.cat([sp.pay_b for sp in sps])
    for sp, s in zip(sps, states):
        sp.gather_b.copy_(gb)
        sp.phase_c(s)
    gc = torch.cat([sp.pay_c for sp in sps])
    res = []
    for sp in sps:
        sp.gather_c.copy_(gc)
        res.append({k: float(v) for k, v in sp.phase_d().items() if k != "scalars"})
    torch.cuda.synchronize()
    assert all(r == res[0] for r in res), res  # bitwise identical scalars on every rank

    osetup = O.set_up_solver(oivp, ts, q)
    ost = O.get_initial_trajectory(osetup)
    odom = O.linearize_at(osetup, ost.mean[1:])
    oout, onll, oobj, ossq, ossqp = O.linear_filtsmooth(osetup["x0"], osetup["dtm"], odom)
    m = torch.cat([s[4] for s in states]).cpu().numpy()
    Lc = torch.cat([s[5] for s in states]).cpu().numpy()
    assert m.shape == oout.mean.shape
    E0 = osetup["E0"]
    y, yo = m @ E0.T, oout.mean @ E0.T
    assert (np.abs(y - yo) <= 1e-9 * np.abs(yo).max(axis=0) + 1e-12).all()
    C, Co = Lc @ np.swapaxes(Lc, -1, -2), oout.chol @ np.swapaxes(oout.chol, -1, -2)
    assert np.abs(C - Co).max() <= 1e-7 * np.abs(Co).max()
    r = res[0]
    assert abs(r["nll"] - onll) <= 1e-9 * abs(onll) + 1e-9 and abs(r["obj"] - oobj) <= 1e-9 * abs(oobj)
    assert abs(r["ssq_proper"] - ossqp) <= 1e-8 * abs(ossqp) and abs(r["ssq"] - ossq) <= 1e-2 * abs(ossq)
