"""Multi-rank time sharding on CPU: world_size 2 and 3 over gloo, stages executed by the host simulator, through the
product's own orchestration (pof.sharded.ShardedPass).  Checks sharded == oracle on every rank's rows."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, q, L, name, ret, family="thread", kw=None):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200"), os.path.join(ROOT, "tests", "hostsim")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from backend import HostBackend
        from oracle import ivps, pof_oracle as O
        from pof.sharded import ShardedPass, shard_bounds

        ivp = getattr(ivps, name)(**(kw or {}))
        ts = np.linspace(ivp.t0, ivp.tmax, N)
        setup = O.set_up_solver(ivp, ts, q)
        st = O.get_initial_trajectory(setup)
        dom = O.linearize_at(setup, st.mean[1:])
        d = setup["d"]
        D = d * (q + 1)
        _, qL = O.preconditioned_discretize_1d(q)
        k_lo, k_hi = shard_bounds(N - 1, rank, world)
        be = HostBackend(d, q, k_hi - k_lo, L, qL, family=family)
        sp = ShardedPass(N, d, q, qL, rank=rank, world=world, device=torch.device("cpu"), backend=be)
        r0 = 0 if rank == 0 else k_lo + 1
        H = torch.from_numpy(np.ascontiguousarray(dom.H[k_lo:k_hi]))
        c = torch.from_numpy(np.ascontiguousarray(dom.b[k_lo:k_hi]))
        means = torch.from_numpy(np.ascontiguousarray(st.mean[r0:k_hi + 1]))
        chols = torch.zeros((sp.rows, D, D), dtype=torch.float64)
        res = sp.run(torch.from_numpy(setup["x0"].mean.copy()), torch.from_numpy(setup["x0"].chol.copy()), H, c, means,
                     chols, calibrate=False)
        out, nll, obj, ssq, ssqp = O.linear_filtsmooth(setup["x0"], setup["dtm"], dom)
        em = np.abs(means.numpy() - out.mean[r0:k_hi + 1]).max() / np.abs(out.mean).max()
        cov = lambda L_: L_ @ np.swapaxes(L_, -1, -2)
        ec = np.abs(cov(chols.numpy()) - cov(out.chol[r0:k_hi + 1])).max() / np.abs(cov(out.chol)).max()
        ok = (em < 1e-9 and ec < 1e-9 and abs(float(res["nll"]) - nll) <= 1e-9 * abs(nll)
              and abs(float(res["obj"]) - obj) <= 1e-9 * abs(obj) and abs(float(res["ssq_proper"]) - ssqp) <= 1e-9 * ssqp
              and abs(float(res["ssq"]) - ssq) <= 1e-2 * ssq)
        ret[rank] = (bool(ok), float(em), float(ec), float(res["nll"]), float(nll))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,N,q,L,name", [(2, 200, 3, 7, "fitzhughnagumo"), (3, 130, 2, 5, "lotkavolterra"),
                                             (2, 64, 3, 4, "logistic")])
def test_sharded_pass_matches_oracle(native_lib, world, N, q, L, name):
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q, L, name, ret)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for r in range(world):
        assert ret[r][0], ret[r]


@pytest.mark.parametrize("world,N,q,L,name,kw", [(2, 120, 3, 7, "fitzhughnagumo", {}),
                                                (3, 70, 2, 5, "lorenz96", {"tmax": 1.0, "d": 8})])
def test_sharded_pass_on_the_tile_family(native_lib, world, N, q, L, name, kw):
    """the same orchestration with the large-state kernels' device code behind the three stages (D = 24 for the
    Lorenz-96 case: beyond the (d <= 4)-templated families)"""
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, N, q, L, name, ret, "tile", kw)) for r in range(world)]
    [p.start() for p in procs]
    [p.join(300) for p in procs]
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    for r in range(world):
        assert ret[r][0], ret[r]
