"""Wall-clock time of full `solve` calls (device-side IEKS loop) at small and medium N: FHN order 3, init="constant"."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np, torch
import pof.ivp
from pof.solver import solve

from pof import _native as nat

ivp = pof.ivp.fitzhughnagumo()
rows = []
for e, graph in [(e, g) for e in [8, 10, 12, 14, 16, 19] for g in (True, False)]:
    nat.USE_LOOP_GRAPH = graph
    ts = np.linspace(0, 100, 2 ** e)
    solve(f=ivp.f, y0=ivp.y0, ts=ts, order=3, init="constant", maxiters=1000)
    torch.cuda.synchronize()
    best, its = 1e9, 0
    for _ in range(3):
        t0 = time.perf_counter()
        ys, info = solve(f=ivp.f, y0=ivp.y0, ts=ts, order=3, init="constant", maxiters=1000)
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
        its = info["iterations"]
    rows.append({"log2n": e, "loop": "while-graph" if graph else "bursts of captured iterations", "iterations": its, "solve_s": best, "ms_per_iteration_incl_setup": 1e3 * best / its})
    print(rows[-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r02_time_solves.json"), "w"), indent=1)
