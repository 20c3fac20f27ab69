import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np, torch
import pof.ivp
from pof import _native as nat
from pof.solver import solve

ivp = pof.ivp.logistic()
ts = np.linspace(ivp.t0, ivp.tmax, 256)
keep = []
for dtype in (torch.float32, torch.float64):
    for graph in (True, False, True):
        nat.USE_LOOP_GRAPH = graph
        ys, info = solve(f=ivp.f, y0=ivp.y0, ts=ts, order=2, init="constant", maxiters=200, dtype=dtype)
        print(dtype, graph, info["iterations"], "sum now", float(ys.mean.sum()), ys.mean.data_ptr(), flush=True)
        torch.cuda.synchronize()
        print("   after sync", float(ys.mean.sum()), "earlier results:", [float(k.sum()) for k in keep], flush=True)
        keep.append(ys.mean)
