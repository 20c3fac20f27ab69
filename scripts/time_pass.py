"""Quick device timing of one IEKS iteration (linearise + filter/smoother pass) at several N."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np
import torch

import pof.ivp
from pof import _native as nat
from pof.convenience import get_initial_trajectory, set_up_solver
from pof.parallel_filtsmooth import run_pass
from pof.step import linearize_into


def main():
    Ns = [int(a) for a in sys.argv[1:]] or [2**12, 2**16, 2**20]
    ivp = pof.ivp.fitzhughnagumo()
    for N in Ns:
        ts = np.linspace(0, 100, N)
        setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=ts, order=3)
        st = get_initial_trajectory(setup, method="constant")
        lin = setup["om"].f._pof_lin
        d, q, D = 2, 3, 8
        dev = setup["_device"]
        means0 = st.mean.contiguous()
        means = means0.clone()
        chols = torch.empty((N, D, D), dtype=torch.float64, device=dev)
        H = torch.empty((N - 1, d, D), dtype=torch.float64, device=dev)
        c = torch.empty((N - 1, d), dtype=torch.float64, device=dev)
        sc = torch.zeros(nat.NSCALARS, dtype=torch.float64, device=dev)
        for L in [int(x) for x in os.environ.get('CHUNK_LENS', '0').split(',')]:
            L = L or None
            def it():
                linearize_into(lin, means, H, c)
                run_pass(setup["x0"], setup["_qL"], H, c, means, chols, d=d, q=q, calibrate=True, chunk_len=L,
                         scalars=sc)
            for _ in range(2):
                means.copy_(means0); it()
            torch.cuda.synchronize()
            tms = []
            for _ in range(3):
                means.copy_(means0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); it(); e1.record(); torch.cuda.synchronize()
                tms.append(e0.elapsed_time(e1))
            print(f"N={N} L={L or nat.default_chunk_len(N, d, q)} ms/iter={min(tms):.3f} scalars={sc.cpu().numpy()[:5]}", flush=True)


if __name__ == "__main__":
    main()
