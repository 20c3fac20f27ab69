"""Regularised iteration on a problem where the plain IEKS diverges: Henon-Heiles (tmax = 10),
order 3, N = 2^17 -- the reference's published run hits maxiters with error 1e6..1e10 (henonheiles_*.csv), and so does
this implementation's plain `solve`.  The quadratic-penalty iteration (`pof.iterators.qpm_ieks_iterator`, reference
iterators.py:53-112) runs every pass with observation noise (reg / n) I on the large-state CUDA kernels.

    python scripts/qpm_henonheiles.py [--log2n 17] [--order 3]
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
from scipy.integrate import solve_ivp  # noqa: E402

import pof.ivp  # noqa: E402
from pof.iterators import qpm_ieks_iterator  # noqa: E402
from pof.solver import solve  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--log2n", type=int, default=17)
ap.add_argument("--order", type=int, default=3)
ap.add_argument("--max-passes", type=int, default=600)
a = ap.parse_args()
N = 2 ** a.log2n
ivp = pof.ivp.henonheiles(tmax=10.0)
ts = np.linspace(ivp.t0, ivp.tmax, N)
y0 = ivp.y0.cpu().numpy()


def rhs(t, y, p=1.0):  # pof/ivp.py:137-152 (truth for the error only)
    return [y[2], y[3], -y[0] - 2 * p * y[0] * y[1], -y[1] - p * (y[0] ** 2 - y[1] ** 2)]


ref = solve_ivp(rhs, (ivp.t0, ivp.tmax), y0, method="DOP853", rtol=1e-13, atol=1e-13, t_eval=ts).y.T
t0 = time.perf_counter()
ys, info = solve(f=ivp.f, y0=ivp.y0, ts=ts, order=a.order, init="constant", maxiters=300)
torch.cuda.synchronize()
t_plain = time.perf_counter() - t0
y = ys.mean.cpu().numpy()
ok = np.isfinite(y).all(axis=1)
rmse_plain = float(np.linalg.norm(y[ok] - ref[ok], axis=1).mean()) if ok.any() else float("nan")
print(f"plain IEKS: {info['iterations']} iterations, rmse_traj {rmse_plain:.3e}, {t_plain:.2f} s", flush=True)

t0 = time.perf_counter()
it, setup = qpm_ieks_iterator(f=ivp.f, y0=ivp.y0, ts=ts, order=a.order, init="constant")
k = 0
for st, nll, obj, reg in it:
    k += 1
    if k >= a.max_passes:
        break
torch.cuda.synchronize()
t_qpm = time.perf_counter() - t0
yq = (st.mean @ setup["E0"].T).cpu().numpy()
ok = np.isfinite(yq).all(axis=1)
rmse_qpm = float(np.linalg.norm(yq[ok] - ref[ok], axis=1).mean()) if ok.any() else float("nan")
print(f"QPM IEKS: {k} passes, final reg {reg:g}, rmse_traj {rmse_qpm:.3e}, {t_qpm:.2f} s", flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump({"problem": "henonheiles(tmax=10)", "order": a.order, "N": N,
           "plain": {"iterations": info["iterations"], "rmse_traj": rmse_plain, "seconds": t_plain},
           "qpm": {"passes": k, "final_reg": reg, "rmse_traj": rmse_qpm, "seconds": t_qpm}},
          open(os.path.join(ROOT, "gpurun_out", f"r02_qpm_henonheiles_n{a.log2n}.json"), "w"), indent=1)
