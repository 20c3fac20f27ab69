"""ms per IEKS iteration vs N (BASELINE config 2: N = 2^6 .. 2^20, FHN order 3, one B200, fp64): the fused iteration
replayed from its CUDA graph, plus per-segment device times of an eager pass.

    python scripts/sweep_n.py [--exps 6,10,14,20] [--flags F] [--tag name] [--d 2 --order 3 --ivp fitzhughnagumo]
flags: 4 = one launch per tree level (A/B against the dataflow sweeps), 1 = large-state kernel family
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import pof.ivp  # noqa: E402
from pof import _native as nat  # noqa: E402
from pof.convenience import get_initial_trajectory, set_up_solver  # noqa: E402
from pof.parallel_filtsmooth import GraphedIteration  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--exps", default=",".join(str(e) for e in range(6, 21)))
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--tag", default="sweep")
ap.add_argument("--ivp", default="fitzhughnagumo")
ap.add_argument("--order", type=int, default=3)
a = ap.parse_args()
nat.DEFAULT_FLAGS = a.flags
ivp = getattr(pof.ivp, a.ivp)()
rows = []
for e in [int(x) for x in a.exps.split(",")]:
    N = 2 ** e
    ts = np.linspace(ivp.t0, ivp.tmax, N)
    setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=ts, order=a.order)
    lin = setup["om"].f._pof_lin
    means0 = get_initial_trajectory(setup, method="constant", means_only=True).mean.contiguous()
    means = means0.clone()
    D = means.shape[1]
    chols = torch.empty((N, D, D), dtype=torch.float64, device=means.device)
    sc = torch.zeros(nat.NSCALARS, dtype=torch.float64, device=means.device)
    it = GraphedIteration(setup["x0"], setup["_qL"], lin, means, chols, sc, calibrate=True)
    for _ in range(3):
        it()
    torch.cuda.synchronize()
    it.ws.ctx.profile_enable(True)
    for _ in range(3):
        it()
    seg = it.ws.ctx.profile_read()
    it.ws.ctx.profile_enable(False)
    it.capture()
    reps = 20 if N <= 2 ** 16 else 5
    best = 1e9
    for _ in range(3):
        means.copy_(means0)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            it()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    L = it.ws.chunk_len
    launches = 1 + int(nat.LIB.pof_launches_per_pass(N, lin["d"], lin["q"], L, nat.flags()))
    segs = {k: round(v[0] / max(1, v[1]), 5) for k, v in seg.items()}
    rows.append({"log2n": e, "N": N, "chunk_len": L, "ms_per_iteration": best, "launches": launches,
                 "segments_ms_eager": segs, "finite": bool(torch.isfinite(sc[:4]).all().item())})
    print(f"N=2^{e} L={L} launches={launches} ms/iter={best:.4f} segs={segs}", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"flags": a.flags, "ivp": a.ivp, "order": a.order, "rows": rows},
          open(os.path.join(ROOT, "gpurun_out", a.tag + ".json"), "w"), indent=1)
