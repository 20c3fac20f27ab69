"""A few fused IEKS iterations (compact linearisation: the bench path) for ncu captures.
    ncu --set full -k regex:k_lane2 -c 3 ... python scripts/profile_iter.py [N] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np
import torch

import pof.ivp
from pof import _native as nat
from pof.convenience import get_initial_trajectory, set_up_solver
from pof.parallel_filtsmooth import run_iteration

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2**20
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ivp = pof.ivp.fitzhughnagumo()
setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=np.linspace(0, 100, N), order=3)
st = get_initial_trajectory(setup, method="constant", means_only=True)
lin = setup["om"].f._pof_lin
means = st.mean.contiguous().clone()
chols = torch.empty((N, 8, 8), dtype=torch.float64, device=means.device)
sc = torch.zeros(nat.NSCALARS, dtype=torch.float64, device=means.device)
L = int(os.environ.get("POF_CHUNK_LEN", 0)) or None
for _ in range(iters):
    run_iteration(setup["x0"], setup["_qL"], lin, means, chols, calibrate=True, chunk_len=L, scalars=sc)
torch.cuda.synchronize()
print("scalars", sc.cpu().numpy()[:5])
