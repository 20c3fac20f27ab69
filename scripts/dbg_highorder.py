"""Debug aid: full solves at high order (D > 16: lane1 leaves + generic trees) under the different kernel families."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
    import numpy as np, torch
    import pof.ivp
    from pof.solver import solve
    for name, kw, q, N in [("rigid_body", {}, 5, 1024), ("rigid_body", {}, 5, 128), ("henonheiles", {"tmax": 10.0}, 5, 2048),
                           ("rigid_body", {}, 4, 1024), ("lotkavolterra", {}, 5, 8192)]:
        ivp = getattr(pof.ivp, name)(**kw)
        ts = np.linspace(ivp.t0, ivp.tmax, N)
        ys, info = solve(f=ivp.f, y0=ivp.y0, ts=ts, order=q, init="constant", maxiters=60)
        print(f"  {name} q={q} N={N}: iterations={info['iterations']} finite={bool(torch.isfinite(ys.mean).all())} "
              f"nll={info['nll']:.6e} obj={info['obj']:.6e}", flush=True)
else:
    for env in [{}, {"POF_B200_LEAF_IMPL": "thread"}, {"POF_B200_LEAF_IMPL": "lane1"}, {"POF_B200_TREE_IMPL": "generic"},
                {"POF_B200_TREE_APEX": "0"}, {"POF_B200_OVERLAP": "0"}]:
        print("env", env, flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), timeout=600)
