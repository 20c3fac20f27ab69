"""Full solves at large N against the reference's published iteration counts / errors (FHN, order 3)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np, torch
from scipy.integrate import solve_ivp
import pof.ivp
from pof.solver import solve
from oracle import ivps as oivps

pub = {r["N"]: r for r in json.load(open(os.path.join(ROOT, "tests/golden/published_ieks3.json")))["fitzhughnagumo"]}
ivp = pof.ivp.fitzhughnagumo(); oivp = oivps.fitzhughnagumo()
for N in [int(a) for a in sys.argv[1:]] or [8192, 65536]:
    ts = np.linspace(0, 100, N)
    t0 = time.time()
    ys, info = solve(f=ivp.f, y0=ivp.y0, ts=ts, order=3, init="constant", maxiters=1000)
    torch.cuda.synchronize(); dt = time.time() - t0
    sol = solve_ivp(oivp.f, (0, 100), oivp.y0, method="DOP853", rtol=1e-13, atol=1e-13, t_eval=ts)
    rmse = np.mean(np.linalg.norm(ys.mean.cpu().numpy() - sol.y.T, axis=1))
    p = pub.get(N, {})
    print(f"N={N} iterations={info['iterations']} (published {p.get('IEKS(3)_iterations')}) rmse={rmse:.4e} (published {p.get('IEKS(3)_rmse_traj')}) "
          f"runtime={dt:.3f}s (published V100 {p.get('IEKS(3)_runtime')}) nll={info['nll']:.6e} obj={info['obj']:.6e} ssq={info['sigma_squared']:.4e}", flush=True)
