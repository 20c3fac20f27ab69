"""ms per IEKS iteration vs N (BASELINE config 2: N = 2^6 .. 2^20, FHN order 3, one B200, fp64)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np, torch
import pof.ivp
from pof import _native as nat
from pof.convenience import get_initial_trajectory, set_up_solver
from pof.parallel_filtsmooth import run_iteration

ivp = pof.ivp.fitzhughnagumo()
rows = []
use_graph = "--graph" in sys.argv
for e in range(6, 21):
    N = 2 ** e
    ts = np.linspace(0, 100, N)
    setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=ts, order=3)
    st = get_initial_trajectory(setup, method="constant")
    lin = setup["om"].f._pof_lin
    means0 = st.mean.contiguous(); means = means0.clone()
    chols = torch.empty((N, 8, 8), dtype=torch.float64, device=means.device)
    sc = torch.zeros(nat.NSCALARS, dtype=torch.float64, device=means.device)
    it = lambda: run_iteration(setup["x0"], setup["_qL"], lin, means, chols, calibrate=True, scalars=sc)
    for _ in range(3): it()
    torch.cuda.synchronize()
    g = None
    if use_graph:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            it(); torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                it()
        torch.cuda.synchronize()
    reps = 20 if N <= 2**16 else 5
    best = 1e9
    for _ in range(3):
        means.copy_(means0); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            g.replay() if g else it()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    L = nat.default_chunk_len(N, 2, 3)
    rows.append((N, L, best))
    print(f"N=2^{e}={N} L={L} ms/iter={best:.4f}", flush=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "sweep_graph.json" if use_graph else "sweep.json"), "w"))
