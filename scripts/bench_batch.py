"""Batch-of-IVPs axis: B independent FitzHugh-Nagumo problems (perturbed initial values), N points
each, solved to convergence -- `pof.batch.solve_batch` (lockstep on separate streams, one host sync per round) against
B consecutive `pof.solver.solve` calls.  Below N ~ 2^16 one IEKS iteration is the latency of its tree sweeps and uses a
fraction of the GPU; the batch fills it.

    python scripts/bench_batch.py [--log2n 10] [--batches 1,8,32,64]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import pof.ivp  # noqa: E402
from pof.batch import solve_batch  # noqa: E402
from pof.solver import solve  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=10)
ap.add_argument("--batches", default="1,8,32,64")
a = ap.parse_args()
N = 2 ** a.log2n
ivp = pof.ivp.fitzhughnagumo()
ts = np.linspace(ivp.t0, ivp.tmax, N)
rows = []
for B in [int(x) for x in a.batches.split(",")]:
    probs = [dict(f=ivp.f, y0=ivp.y0 * (1.0 + 0.002 * i), ts=ts) for i in range(B)]
    res = solve_batch(probs, order=3, init="constant", maxiters=1000)  # warm-up (allocations, kernel attributes)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = solve_batch(probs, order=3, init="constant", maxiters=1000)
    torch.cuda.synchronize()
    t_batch = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = [solve(f=p["f"], y0=p["y0"], ts=p["ts"], order=3, init="constant", maxiters=1000) for p in probs]
    torch.cuda.synchronize()
    t_seq = time.perf_counter() - t0
    same = all(torch.equal(r[0].mean, s[0].mean) for r, s in zip(res, ref))
    its = sum(r[1]["iterations"] for r in res)
    rows.append({"N": N, "batch": B, "solve_batch_s": t_batch, "consecutive_solves_s": t_seq, "speedup": t_seq / t_batch,
                 "iterations_total": its, "identical_results": bool(same),
                 "ms_per_iteration_per_problem_batched": 1e3 * t_batch / its})
    print(rows[-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", f"r02_batch_n{a.log2n}.json"), "w"), indent=1)
