"""BASELINE config 5: larger-state IVP (Lorenz-96, d = 16, q = 3, D = 64) at N = 2^18, one B200, fp64 -- ms per IEKS
iteration on the large-state ("tile") kernels, per-segment device times, and the FP64 roofline fraction computed from
the reference-formula FLOP count of SURVEY.md 8d.  Standalone (bench.py's contract line stays config 2).

    python scripts/bench_config5.py [--log2n 18] [--steps 3] [--warmup 1]
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

import pof.ivp  # noqa: E402
from pof import _native as nat  # noqa: E402
from pof.convenience import get_initial_trajectory, set_up_solver  # noqa: E402
from pof.parallel_filtsmooth import run_iteration  # noqa: E402


def flop_step(D, d):
    """reference-formula FLOPs per time step per iteration"""
    cf = 121 / 3 * D**3 + 12 * D**2
    cs = 22 / 3 * D**3 + 2 * D**2
    ef = 22 / 3 * D**3 + 4 / 3 * (D + d) ** 3 + 6 * d * D**2 + 3 * d**2 * D + 6 * d * D
    es = 46 / 3 * D**3 + 4 * D**2
    o = 16 / 3 * D**3 + 2 * D**2 + 2 * d * D**2 + 2 * (D + d) * d**2 - 2 / 3 * d**3 + 2 * d * D + 2 * d**2
    misc = 8 * D**2 + 4 * d * D
    return 2 * cf + 2 * cs + ef + es + o + misc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2n", type=int, default=18)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--d", type=int, default=16)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--flags", type=int, default=0, help="2 = shared-memory Householder sweeps (A/B)")
    a = ap.parse_args()
    nat.DEFAULT_FLAGS = a.flags
    torch.cuda.set_device(0)
    d, q = a.d, a.order
    D = d * (q + 1)
    N = 2**a.log2n
    ivp = pof.ivp.lorenz96(d=d, tmax=10.0)
    ts = np.linspace(ivp.t0, ivp.tmax, N)
    setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=ts, order=q)
    st = get_initial_trajectory(setup, method="constant", means_only=True)
    lin = setup["om"].f._pof_lin
    means0 = st.mean.contiguous()
    means = means0.clone()
    chols = torch.empty((N, D, D), dtype=torch.float64, device=means.device)
    sc = torch.zeros(nat.NSCALARS, dtype=torch.float64, device=means.device)
    L = nat.default_chunk_len(N, d, q)
    ws = nat.Workspace(N, d, q, L, means.device)
    it = lambda: run_iteration(setup["x0"], setup["_qL"], lin, means, chols, calibrate=True, scalars=sc, chunk_len=L,
                               ws=ws)
    for _ in range(max(a.warmup, 1)):
        it()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        it()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    # per-segment device times (CUDA events inside the library)
    ws.ctx.profile_enable(True)
    it()
    segd = ws.ctx.profile_read()
    ws.ctx.profile_enable(False)
    peak = ctypes.c_double(0.0)
    nat.LIB.pof_measure_dfma_tflops(nat.stream_ptr(), ctypes.byref(peak))
    flops = flop_step(D, d) * N
    achieved = flops / (ms * 1e-3) / 1e12
    out = {
        "metric": "ms per IEKS iteration (fp64)", "value": ms, "unit": "ms", "n_gpus": 1, "steps": a.steps,
        "warmup": a.warmup, "higher_is_better": False, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"BASELINE config 5: Lorenz-96 d={d} order={q} (D={D}), N=2^{a.log2n}, constant init",
                   "chunk_len": L, "kernels": "tile (CTA per chunk, shared-memory tiles)"},
        "segments_ms": {k: v[0] for k, v in segd.items()},
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak.value, "unit": "TFLOP/s",
                     "frac": achieved / peak.value if peak.value else None,
                     "note": "reference-formula FLOPs (SURVEY 8d); the leaf recursions execute ~8x fewer"},
        "finite": bool(torch.isfinite(means).all().item()), "nll": float(sc[nat.S_NLL]), "obj": float(sc[nat.S_OBJ]),
    }
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bench_config5.json"), "w"))


if __name__ == "__main__":
    main()
