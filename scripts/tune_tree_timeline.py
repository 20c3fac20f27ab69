"""Tuning run (library built with POF_NVCC_EXTRA=-DPOF_TUNE): per-level completion times of the filter sweep.  In tuning
builds every ready flag holds the global-timer value at which its node completed; one eager pass, then the flag area
is read back."""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np, torch
import pof.ivp
from pof import _native as nat
from pof.convenience import get_initial_trajectory, set_up_solver
from pof.parallel_filtsmooth import run_iteration

lib = nat.LIB
lib.pof_tune_ks_max.argtypes = [ctypes.c_long]
lib.pof_tune_flag_layout.argtypes = [ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_long, ctypes.POINTER(ctypes.c_long)]
ivp = pof.ivp.fitzhughnagumo()
out = {}
for e in (10, 20):
    N = 2 ** e
    setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=np.linspace(0, 100, N), order=3)
    lin = setup["om"].f._pof_lin
    for ks_max in (1536, 1):
        lib.pof_tune_ks_max(ks_max)
        L = nat.default_chunk_len(N, 2, 3, 0)
        lay = (ctypes.c_long * 64)()
        lib.pof_tune_flag_layout(N - 1, 2, 3, L, lay)
        off, total, nlev, B, K, nB, ksw = [int(lay[i]) for i in range(7)]
        sz = [int(lay[7 + l]) for l in range(nlev)]
        lev_off = np.concatenate([[0], np.cumsum(sz)])
        means = get_initial_trajectory(setup, method="constant", means_only=True).mean.contiguous()
        chols = torch.empty((N, 8, 8), dtype=torch.float64, device=means.device)
        ws = nat.Workspace(N, 2, 3, L, means.device, torch.float64)
        for _ in range(3):
            run_iteration(setup["x0"], setup["_qL"], lin, means, chols, calibrate=True, chunk_len=L, ws=ws)
        torch.cuda.synchronize()
        fl = ws.buf[off:off + 4 * (ksw + K * nB)].view(torch.int32).cpu().numpy().astype(np.int64) & 0x7fffffff
        f_up, f_dn = fl[64:64 + total], fl[64 + total:64 + 2 * total]
        ks = fl[ksw:ksw + K * nB].reshape(K, nB) if K else np.zeros((0, 0), dtype=np.int64)
        ev = []
        for l in range(1, nlev):
            v = f_up[lev_off[l]:lev_off[l + 1]]
            v = v[v > 0]
            if v.size:
                ev.append(("up%d" % l, int(v.min()), int(v.max()), int(v.size)))
        for s_ in range(K):
            v = ks[s_][ks[s_] > 0]
            if v.size:
                ev.append(("ks%d" % (s_ + 1), int(v.min()), int(v.max()), int(v.size)))
        for l in range(nlev - 1, -1, -1):
            v = f_dn[lev_off[l]:lev_off[l + 1]]
            v = v[v > 0]
            if v.size:
                ev.append(("dn%d" % l, int(v.min()), int(v.max()), int(v.size)))
        t0 = min(a for _, a, _, _ in ev)
        rows = [(nm, (a - t0) / 1e3, (b - t0) / 1e3, c) for nm, a, b, c in ev]
        out["n%d_ks%d" % (e, ks_max)] = rows
        print("N=2^%d ks_max=%d  B=%d K=%d nB=%d  (first, last completion in us; count)" % (e, ks_max, B, K, nB))
        for r in rows:
            print("   %-6s %8.1f %8.1f %6d" % r)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02t_tree_timeline.json"), "w"), indent=1)
