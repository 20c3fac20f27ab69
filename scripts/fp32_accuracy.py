"""Where the optional fp32 mode is usable: fp32 vs fp64 GPU results of ONE IEKS iteration (FitzHugh-Nagumo, order 3) as
N grows, from the constant initial trajectory and from the fp64-converged trajectory, plus the time per iteration of
both modes.  Writes gpurun_out/r02_fp32_accuracy.md (copied to profiles/)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import pof.ivp  # noqa: E402
from pof import _native as nat  # noqa: E402
from pof.convenience import get_initial_trajectory, set_up_solver  # noqa: E402
from pof.convergence_criteria import crit_scalars  # noqa: E402
from pof.parallel_filtsmooth import GraphedIteration, run_iteration  # noqa: E402
from pof.utils import MVNSqrt  # noqa: E402

ivp = pof.ivp.fitzhughnagumo()
rows = []
for e in [6, 8, 10, 12, 14, 16, 18, 20]:
    N = 2 ** e
    ts = np.linspace(0, 100, N)
    setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=ts, order=3)
    lin = setup["om"].f._pof_lin
    E0 = setup["E0"]
    m0 = get_initial_trajectory(setup, method="constant", means_only=True).mean.contiguous()
    conv = m0.clone()
    ch = torch.empty((N, 8, 8), dtype=torch.float64, device=m0.device)
    obj_old = nll_old = 0.0
    for k in range(400):
        sc = run_iteration(setup["x0"], setup["_qL"], lin, conv, ch, calibrate=True).cpu()
        if k >= 1 and crit_scalars(float(sc[1]), obj_old, float(sc[0]), nll_old, float(sc[4])):
            break
        nll_old, obj_old = float(sc[0]), float(sc[1])
    res = {}
    for start, mm in (("constant", m0), ("converged", conv)):
        a = mm.clone()
        ca = torch.empty((N, 8, 8), dtype=torch.float64, device=a.device)
        run_iteration(setup["x0"], setup["_qL"], lin, a, ca, calibrate=False)
        b = mm.to(torch.float32).contiguous()
        cb = torch.empty((N, 8, 8), dtype=torch.float32, device=a.device)
        x032 = MVNSqrt(setup["x0"].mean.float(), setup["x0"].chol.float())
        run_iteration(x032, setup["_qL"], lin, b, cb, calibrate=False)
        ya, yb = a @ E0.T, b.double() @ E0.T
        res[start] = float(((ya - yb).abs().max(dim=0).values / ya.abs().max(dim=0).values).max())
    times = {}
    for dt in (torch.float64, torch.float32):
        mm = m0.to(dt).contiguous()
        cc = torch.empty((N, 8, 8), dtype=dt, device=mm.device)
        sc = torch.zeros(nat.NSCALARS, dtype=dt, device=mm.device)
        x0 = MVNSqrt(setup["x0"].mean.to(dt), setup["x0"].chol.to(dt))
        it = GraphedIteration(x0, setup["_qL"], lin, mm, cc, sc, calibrate=True)
        for _ in range(3):
            it()
        it.capture()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            it()
        e1.record()
        torch.cuda.synchronize()
        times[dt] = e0.elapsed_time(e1) / 5
    rows.append((e, res["constant"], res["converged"], times[torch.float64], times[torch.float32]))
    print(rows[-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "r02_fp32_accuracy.md"), "w") as fh:
    fh.write("# fp32 mode: accuracy and speed (FitzHugh-Nagumo, order 3, one IEKS iteration, B200)\n\n"
             "max over components of |E0 m (fp32) - E0 m (fp64)| / max|E0 m|, fp64 = this library's fp64 kernels on "
             "the same input\n\n| N | from the constant trajectory | from the converged trajectory | fp64 ms/iter | "
             "fp32 ms/iter |\n|---|---|---|---|---|\n")
    for e, a, b, t64, t32 in rows:
        fh.write(f"| 2^{e} | {a:.2e} | {b:.2e} | {t64:.3f} | {t32:.3f} |\n")
    ok = [e for e, a, b, *_ in rows if a <= 1e-4]
    fh.write(f"\nLargest N with <= 1e-4 on the outputs from the constant trajectory: 2^{max(ok) if ok else '-'}\n")
