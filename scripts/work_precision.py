"""Work-precision runner in the shape of the reference's experiments/3_work_precision_diagram/run_benchmark.py
(BASELINE config 3): IEKS(q) parallel, sIEKS(q) (solve(sequential=True)) and EKS(q) on the pof.ivp benchmark problems,
`solve(init="constant", maxiters=1000)` on ts = linspace(t0, tmax, N); columns named as in the reference's CSVs
(`Ns`, `<method>_runtime`, `<method>_rmse_final`, `<method>_rmse_traj`, `<method>_iterations`) plus, where the
reference published a value (tests/golden/published_ieks3.json: its V100 CSVs), `<method>_iterations_published` and
`<method>_rmse_traj_published`.  The reference trajectory is SciPy DOP853 at rtol = atol = 1e-13 (the reference used
diffrax Kvaerno5 at rtol 1e-13, not available here).

    python scripts/work_precision.py [--setups logistic,fhn] [--orders 1,2,3] [--max-exp 19] [--out profiles]
"""
import argparse
import csv
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np
import torch
from scipy.integrate import solve_ivp

import pof.ivp
from pof.solver import sequential_eks_solve, solve

from oracle import ivps as oivps  # vector fields for the SciPy reference trajectory only

SETUPS = {  # name -> (factory name, kwargs, first exponent, key in the published fixture)
    "logistic": ("logistic", {}, 4, "logistic"),
    "fhn": ("fitzhughnagumo", {}, 7, "fitzhughnagumo"),
    "lotkavolterra": ("lotkavolterra", {}, 7, None),
    "vdp0": ("vanderpol", {"stiffness_constant": 1.0}, 7, "vanderpol_mu1"),
    "rigidbody": ("rigid_body", {}, 7, "rigid_body"),
    "henonheiles": ("henonheiles", {"tmax": 10.0}, 5, "henonheiles_tmax10"),
}


def timed(fn, reps):
    out = fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return out, best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--setups", default=",".join(SETUPS))
    ap.add_argument("--orders", default="1,2,3,4,5")
    ap.add_argument("--max-exp", type=int, default=19)
    ap.add_argument("--exp-step", type=int, default=2)
    ap.add_argument("--seq-max-exp", type=int, default=12, help="largest N for the sequential solvers (one thread)")
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--tag", default="r02")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles"))
    a = ap.parse_args()
    orders = [int(x) for x in a.orders.split(",")]
    published = json.load(open(os.path.join(ROOT, "tests", "golden", "published_ieks3.json")))
    dev = torch.cuda.get_device_name(0).replace(" ", "_")
    for name in a.setups.split(","):
        fac, kw, e0, pubkey = SETUPS[name]
        ivp = getattr(pof.ivp, fac)(**kw)
        oivp = getattr(oivps, fac)(**kw)
        pub = {int(r["N"]): r for r in published.get(pubkey, [])} if pubkey else {}
        Ns = [2**e for e in range(e0, a.max_exp + 1, a.exp_step)]
        ref = solve_ivp(lambda t, y: oivp.f(t, y), (ivp.t0, ivp.tmax), np.asarray(oivp.y0, dtype=float), method="DOP853",
                        rtol=1e-13, atol=1e-13, dense_output=True)
        rows = []
        for N in Ns:
            ts = np.linspace(ivp.t0, ivp.tmax, N)
            yref = ref.sol(ts).T
            row = {"Ns": N}

            def record(method, ys, info, rt):
                y = ys.mean.cpu().numpy()
                ok = np.isfinite(y[:, 0])
                row[f"{method}_runtime"] = rt
                row[f"{method}_rmse_final"] = float(np.linalg.norm(y[ok][-1] - yref[ok][-1])) if ok.any() else np.nan
                row[f"{method}_rmse_traj"] = float(np.linalg.norm(y[ok] - yref[ok], axis=1).mean()) if ok.any() else np.nan
                if "iterations" in info:
                    row[f"{method}_iterations"] = info["iterations"]
                p = pub.get(N, {})
                for k in ("iterations", "rmse_traj", "runtime"):
                    if f"{method}_{k}" in p:
                        row[f"{method}_{k}_published"] = p[f"{method}_{k}"]

            for q in orders:
                (ys, info), rt = timed(lambda: solve(f=ivp.f, y0=ivp.y0, ts=ts, order=q, init="constant",
                                                      maxiters=1000), a.reps)
                record(f"IEKS({q})", ys, info, rt)
                if N <= 2**a.seq_max_exp:
                    (ys, info), rt = timed(lambda: solve(f=ivp.f, y0=ivp.y0, ts=ts, order=q, init="constant",
                                                          maxiters=1000, sequential=True), a.reps)
                    record(f"sIEKS({q})", ys, info, rt)
                    (ys, info), rt = timed(lambda: sequential_eks_solve(f=ivp.f, y0=ivp.y0, ts=ts, order=q), a.reps)
                    record(f"EKS({q})", ys, info, rt)
            rows.append(row)
            print(name, {k: (f"{v:.3e}" if isinstance(v, float) else v) for k, v in row.items()
                         if k == "Ns" or k.startswith("IEKS(3)")}, flush=True)
        cols = ["Ns"] + sorted({k for r in rows for k in r if k != "Ns"})
        path = os.path.join(a.out, f"{a.tag}_work_precision_{name}_{dev}.csv")
        with open(path, "w", newline="") as fh:
            w = csv.DictWriter(fh, fieldnames=cols)
            w.writeheader()
            for r in rows:
                w.writerow(r)
        print("wrote", path, flush=True)


if __name__ == "__main__":
    main()
