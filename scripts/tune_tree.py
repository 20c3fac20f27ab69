"""Tuning run for the hybrid filter sweep (needs a library built with POF_NVCC_EXTRA=-DPOF_TUNE): ms per iteration for
different base levels of the Kogge-Stone stage (WsLayout::KS_MAX) and flag-poll back-offs."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np, torch
import pof.ivp
from pof import _native as nat
from pof.convenience import get_initial_trajectory, set_up_solver
from pof.parallel_filtsmooth import GraphedIteration

lib = nat.LIB
lib.pof_tune_ks_max.argtypes = [ctypes.c_long]
lib.pof_tune_poll_ns.argtypes = [ctypes.c_int]
ivp = pof.ivp.fitzhughnagumo()
rows = []
for e in (10, 16, 19, 20):
    N = 2 ** e
    setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=np.linspace(0, 100, N), order=3)
    lin = setup["om"].f._pof_lin
    for ks_max, poll in [(1, 64), (192, 64), (384, 64), (768, 64), (1536, 64), (3072, 64), (6144, 64), (1536, 16),
                         (1536, 0), (768, 0), (1, 0)]:
        lib.pof_tune_ks_max(ks_max)
        lib.pof_tune_poll_ns(poll)
        means = get_initial_trajectory(setup, method="constant", means_only=True).mean.contiguous()
        chols = torch.empty((N, 8, 8), dtype=torch.float64, device=means.device)
        scalars = torch.zeros(nat.NSCALARS, dtype=torch.float64, device=means.device)
        it = GraphedIteration(setup["x0"], setup["_qL"], lin, means, chols, scalars)
        it(); it(); it.capture()
        for _ in range(5):
            it()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 30
        e0.record()
        for _ in range(reps):
            it()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        rows.append({"log2n": e, "ks_max": ks_max, "poll_ns": poll, "ms": ms,
                     "finite": bool(torch.isfinite(means).all())})
        print(rows[-1], flush=True)
        del it
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r02t_tune_tree.json"), "w"), indent=1)
