#!/usr/bin/env python
"""Benchmark of the IEKS hot path: milliseconds per IEKS iteration (fp64), FitzHugh-Nagumo, order 3.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on the host cores (oracle port)

Workload (BASELINE.json configs[1] / configs[3]): one IEKS iteration = linearise at the previous trajectory + the
parallel-in-time square-root filter/smoother pass + calibration + convergence reductions (reference
pof/step.py:33-45), N = 2^20 time points PER GPU on ts = linspace(0, 100, N_total), starting from the constant
initial trajectory and continuing the IEKS loop from there (every timed step is a real iteration).  With N GPUs the time
axis is sharded contiguously (weak scaling: 2^20 points per rank) and the per-shard carry elements are exchanged with
all-gathers (pof/sharded.py).  A step touches > 3 GB of HBM per GPU, far more than the 126 MB L2, so no explicit L2
flush is needed between timed iterations.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]

import numpy as np  # noqa: E402

METRIC = "ms per IEKS iteration (fp64)"
D_, d_, q_ = 8, 2, 3
# algorithmic FLOPs per time step per iteration of the REFERENCE formulas at (D, d) = (8, 2)
FLOP_STEP = 68.5e3
# attribution to the dominant kernel (filter scan: seeded square-root filter = one filtering combine per step, the
# innovation statistics and the smoother-element build): C_f + O + E_s  
FLOP_STEP_SCAN = 21.4e3 + 3.23e3 + 8.1e3
# algorithmic HBM bytes per step of the scan kernel: read the compact linearisation [J_f | c] (6 doubles), write the
# step's backward kernel (g: 8, E: 64, untriangularised noise factor: D x (D-d) = 48 doubles)
BYTES_STEP_SCAN = (6 + 8 + 64 + 48) * 8.0
# smoother: read the backward kernel (120 doubles) and the previous mean (8), write the new mean (8) and the calibrated
# Cholesky factor in the API layout (64)
BYTES_STEP_SMOOTH = (120 + 8 + 8 + 64) * 8.0


def measured_fp64_pipe(kernel_prefix):
    """FP64-pipe utilisation of the kernel from the committed ncu capture (fraction of peak), or None"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        for k, v in t.items():
            if k.startswith(kernel_prefix):
                return v["fp64_pipe_pct"] / 100.0
    except Exception:
        pass
    return None


def measured_ncu(kernel_prefix, field):
    """one field of the committed ncu --set full capture of a kernel (profiles/r01_traffic.json), or None"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        for k, v in t.items():
            if k.startswith(kernel_prefix):
                return v.get(field)
    except Exception:
        pass
    return None


def measured_traffic(kernel_prefix):
    """dram bytes per launch of the scan kernel from the committed ncu --set full capture (profiles/), or None"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        for k, v in t.items():
            if k.startswith(kernel_prefix):
                return v["dram_bytes_read"] + v["dram_bytes_write"]
    except Exception:
        pass
    return None


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--n-time", type=int, default=2**20, help="time points per GPU")
    ap.add_argument("--cpu-sample", type=int, default=2**15, help="N of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU arms
def _threaded_oracle(nthreads):
    """The oracle's batched QR is effectively single-threaded; split the batch over the host cores (numpy's LAPACK
    gufuncs release the GIL)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import pof_oracle as O

    pool = ThreadPoolExecutor(max_workers=nthreads)
    base_tria = O.tria

    def tria(A):
        if A.ndim < 3 or A.shape[0] < 4 * nthreads:
            return base_tria(A)
        parts = np.array_split(np.arange(A.shape[0]), nthreads)
        outs = list(pool.map(lambda idx: base_tria(A[idx[0]:idx[-1] + 1]), [p for p in parts if len(p)]))
        return np.concatenate(outs, axis=0)

    O.tria = tria
    return O


def cpu_iteration_ms(N_sample, nthreads, repeats=1):
    """one oracle IEKS iteration (reference algorithm, JAX association order) at N_sample points, milliseconds"""
    from oracle import ivps

    O = _threaded_oracle(nthreads)
    ivp = ivps.fitzhughnagumo()
    ts = np.linspace(0, 100, N_sample)
    setup = O.set_up_solver(ivp, ts, q_)
    st = O.get_initial_trajectory(setup)
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        st2, *_ = O.ieks_step(setup, st)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    return best, st2


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    N_s = min(args.cpu_sample, 2**13)
    from oracle import ivps

    O = _threaded_oracle(cores)
    ivp = ivps.fitzhughnagumo()
    ts = np.linspace(0, 100, N_s)
    setup = O.set_up_solver(ivp, ts, q_)
    st = O.get_initial_trajectory(setup)
    for _ in range(args.warmup):
        st, *_ = O.ieks_step(setup, st)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st, *_ = O.ieks_step(setup, st)
    ms_sample = (time.perf_counter() - t0) * 1e3 / args.steps
    N_total = args.n_time * args.gpus
    scale = N_total / N_s
    value = ms_sample * scale
    sample = (f"oracle port of the reference IEKS iteration (NumPy/LAPACK, JAX association order), {cores} threads, "
              f"N={N_s} per step, linearly extrapolated x{scale:g} to N={N_total}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": value, "higher_is_better": False, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "ms", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args):
    return {
        "workload": f"FitzHugh-Nagumo, IWP order 3 (D=8, d=2), one IEKS iteration, N=2^{int(np.log2(args.n_time))} "
                    f"time points per GPU, ts=linspace(0,100,N_total), constant initial trajectory, IEKS loop continued",
        "n_time_per_gpu": args.n_time, "n_time_total": args.n_time * args.gpus, "order": q_, "state_dim": D_,
        "sharding": "contiguous time shards, carry all-gather" if args.gpus > 1 else "single GPU",
        "l2": "working set per step (>3 GB) exceeds L2 (126 MB); no explicit flush",
    }


# ------------------------------------------------------------------------------------------------ native arm
def run_native(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # keep stdout clean for the ONE JSON line: libraries (e.g. NCCL's version banner) print to fd 1
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(obj), flush=True)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    import pof.ivp
    from pof import _native as nat
    from pof.convenience import set_up_solver
    from pof.parallel_filtsmooth import GraphedCall, GraphedIteration, run_iteration
    from pof.sharded import ShardedPass, shard_bounds
    from pof.step import linearize_into

    N_total = args.n_time * world
    n = N_total - 1
    ivp = pof.ivp.fitzhughnagumo()
    # set_up_solver only needs the grid spacing and y0: give it a 2-point grid with the right dt (host, O(1))
    dt = 100.0 / (N_total - 1)
    setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=np.array([0.0, dt]), order=q_)
    lin = setup["om"].f._pof_lin
    x0, qL = setup["x0"], setup["_qL"]
    k_lo, k_hi = shard_bounds(n, rank, world)
    n_loc = k_hi - k_lo
    rows = n_loc + (1 if rank == 0 else 0)
    # constant initial trajectory (reference initialization.py:42-56 then PI @ .)
    y0 = ivp.y0.to(dev)
    row = torch.zeros(D_, dtype=torch.float64, device=dev)
    f0 = ivp.f(None, y0)
    for b in range(d_):
        row[b * (q_ + 1)] = y0[b]
        row[b * (q_ + 1) + 1] = f0[b]
    row = setup["PI"] @ row
    means0 = row.repeat(rows, 1).contiguous()
    means = means0.clone()
    chols = torch.empty((rows, D_, D_), dtype=torch.float64, device=dev)
    if world > 1:
        Jc = torch.empty((n_loc, d_ * d_ + d_), dtype=torch.float64, device=dev)  # compact linearisation [J_f | c]
    scalars = torch.zeros(nat.NSCALARS, dtype=torch.float64, device=dev)
    t1row = 1 if rank == 0 else 0  # local row of the first linearisation point (state k_lo + 1)

    def linearize():
        ivp_id, params = lin["builtin"]
        ph, pp = nat.host_doubles(list(params) + [0.0])
        nat.check(nat.LIB.pof_linearize_ivp_compact_f64(nat.stream_ptr(), ivp_id, pp, len(params), n_loc, d_, q_,
                                                        lin["scale0"], nat.ptr(means[t1row:]), nat.ptr(Jc)),
                  "linearize")

    if world == 1:
        L = nat.default_chunk_len(N_total, d_, q_, dev.index)
        L = int(os.environ.get("POF_CHUNK_LEN", L))
        launches = 1 + int(nat.LIB.pof_launches_per_pass(N_total, d_, q_, L))

        fused = GraphedIteration(x0, qL, lin, means, chols, scalars, calibrate=True, chunk_len=L)

        def step():  # the fused iteration (linearise + pass), replayed from a CUDA graph once captured
            fused()
            return scalars
    else:
        sp = ShardedPass(N_total, d_, q_, qL, rank=rank, world=world, device=dev)
        sp.backend.set_compact(lin["scale0"], lin["scale1"])
        L = sp.backend.chunk_len
        launches = 1 + int(nat.LIB.pof_launches_per_pass(n_loc + 1, d_, q_, L)) + 2

        def eager_step():
            linearize()
            return sp.run(x0.mean, x0.chol, Jc, None, means, chols, calibrate=True)

        # the whole sharded iteration (kernels, NCCL all-gathers, the small torch ops between the stages) replayed
        # from one CUDA graph per rank; POF_BENCH_SHARDED_GRAPH=0 times the eager launches instead
        fused = GraphedCall(eager_step)

        def step():
            return fused()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    # ---- device-resident timing (value); per-segment CUDA events for the roofline in a second, eager, timed loop
    for _ in range(args.warmup):
        step()
    nat.LIB.pof_profile_enable(1)
    ms_eager = timed(step, max(3, args.steps // 2))
    seg_ms = (np.zeros(7), np.zeros(7, dtype=np.int64))
    nat.check(nat.LIB.pof_profile_read(seg_ms[0].ctypes.data_as(nat._c_dp), seg_ms[1].ctypes.data_as(nat._c_dp)),
              "profile_read")
    nat.LIB.pof_profile_enable(0)
    if world == 1 or os.environ.get("POF_BENCH_SHARDED_GRAPH", "1") != "0":
        barrier()
        fused.capture()
        step()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_iter = timed(step, args.steps)
    clk = clocks.stop() if rank == 0 else None
    last = step()
    torch.cuda.synchronize()
    finite = bool(torch.isfinite(means).all().item())

    # ---- end to end through the public call with HOST buffers: H2D of the previous trajectory means from pinned
    # memory, the iteration, D2H of the projected solution means E0 m (N,d) and the scalars
    h_means = torch.empty((rows, D_), dtype=torch.float64).pin_memory()
    h_means.copy_(means.cpu())
    h_y = torch.empty((rows, d_), dtype=torch.float64).pin_memory()
    h_sc = torch.empty(nat.NSCALARS, dtype=torch.float64).pin_memory()
    ymean = torch.empty((rows, d_), dtype=torch.float64, device=dev)

    def e2e_step():
        means.copy_(h_means, non_blocking=True)
        step()
        nat.check(nat.LIB.pof_project_f64(nat.stream_ptr(), rows, d_, q_, setup["_scale0"], None, nat.ptr(means), None,
                                          nat.ptr(ymean), None), "project")
        h_y.copy_(ymean, non_blocking=True)
        if world == 1:
            h_sc.copy_(scalars, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        e2e_step()
    ms_e2e = timed(e2e_step, max(3, args.steps // 2))
    h2d = rows * D_ * 8
    d2h = rows * d_ * 8 + (nat.NSCALARS * 8 if world == 1 else 0)

    # ---- FP64 peak of this device (no FP64 figure in MEASURED_PEAKS.json)
    tf = np.zeros(1)
    nat.check(nat.LIB.pof_measure_dfma_tflops(nat.stream_ptr(), tf.ctypes.data_as(nat._c_dp)), "dfma peak")
    fp64_peak = float(tf[0])

    def leave():
        # A process group whose NCCL kernels live in captured CUDA graphs does not tear down reliably (the 2-GPU run
        # hung in destroy_process_group after printing its line): flush and leave without the teardown.
        if world > 1:
            torch.cuda.synchronize()
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        leave()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (
        6650.0, "fallback (B200_PROFILING.md)")
    names = ["fold", "filter_up", "filter_down", "scan", "smooth_up", "smooth_down", "smooth"]
    seg = {nm: (seg_ms[0][i] / max(1, seg_ms[1][i])) for i, nm in enumerate(names)}
    scan_ms = seg["scan"]
    scan_tflops = FLOP_STEP_SCAN * n_loc / (scan_ms * 1e-3) / 1e12 if scan_ms > 0 else None
    ms_ref_share = ms_eager  # the segment times were taken in the eager loop
    iter_tflops = FLOP_STEP * n_loc / (ms_iter * 1e-3) / 1e12
    roofline = {
        "kernel": "k_lane2_scan<2,3> (filter scan: seeded square-root filter + backward kernels + innovation "
                  "statistics)",
        "bound": "fp64", "achieved": scan_tflops, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": (scan_tflops / fp64_peak) if scan_tflops else None,
        "traffic": measured_traffic("k_lane2_scan") if args.n_time == 2**20 else None,
        "peak_source": "DFMA loop measured in this run (pof_measure_dfma_tflops); nominal 37 TFLOP/s",
        "algorithmic_flop_per_step": FLOP_STEP_SCAN, "avg_launch_ms": scan_ms,
        "share_of_step": scan_ms / ms_ref_share if ms_ref_share else None,
        "fp64_pipe_utilisation_ncu": measured_fp64_pipe("k_lane2_scan"),
        "lsu_pipe_utilisation_ncu": (measured_ncu("k_lane2_scan", "lsu_pipe_pct") or 0) / 100.0 or None,
        "issue_slot_utilisation_ncu": (measured_ncu("k_lane2_scan", "issue_active_pct") or 0) / 100.0 or None,
        "note": "achieved counts the REFERENCE formulas' flops (SURVEY 8d: one general filtering combine, 21.4 kFLOP, per "
                "step); the kernel reaches the same result with a ~8x cheaper leaf recursion, so frac can exceed 1 -- the "
                "executed-instruction view is fp64_pipe_utilisation_ncu (ncu sm__inst_executed_pipe_fp64, profiles/)",
    }
    smooth_ms = seg["smooth"]
    roofline_smooth = {
        "kernel": "k_lane2_smooth<2,3> (seeded square-root RTS recursion; objective, calibration, convergence count)",
        "bound": "hbm", "achieved": BYTES_STEP_SMOOTH * n_loc / (smooth_ms * 1e-3) / 1e9 if smooth_ms > 0 else None,
        "peak": hbm_peak, "unit": "GB/s",
        "frac": BYTES_STEP_SMOOTH * n_loc / (smooth_ms * 1e-3) / 1e9 / hbm_peak if smooth_ms > 0 else None,
        "traffic": measured_traffic("k_lane2_smooth") if args.n_time == 2**20 else None, "peak_source": hbm_src,
        "algorithmic_bytes_per_step": BYTES_STEP_SMOOTH, "avg_launch_ms": smooth_ms,
        "fp64_pipe_utilisation_ncu": measured_fp64_pipe("k_lane2_smooth"),
        "lsu_pipe_utilisation_ncu": (measured_ncu("k_lane2_smooth", "lsu_pipe_pct") or 0) / 100.0 or None,
    }
    roofline_hbm = {
        "kernel": roofline["kernel"], "bound": "hbm", "achieved": BYTES_STEP_SCAN * n_loc / (scan_ms * 1e-3) / 1e9,
        "peak": hbm_peak, "unit": "GB/s", "frac": BYTES_STEP_SCAN * n_loc / (scan_ms * 1e-3) / 1e9 / hbm_peak,
        "traffic": measured_traffic("k_lane2_scan") if args.n_time == 2**20 else None, "peak_source": hbm_src,
        "algorithmic_bytes_per_step": BYTES_STEP_SCAN,
    }
    roofline_iter = {
        "scope": "whole IEKS iteration (all kernels of a step)", "bound": "fp64", "achieved": iter_tflops,
        "peak": fp64_peak, "unit": "TFLOP/s", "frac": iter_tflops / fp64_peak, "algorithmic_flop_per_step": FLOP_STEP,
        "segments_ms": seg, "ms_per_step_eager_launches": ms_eager,
    }

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ms_s, _ = cpu_iteration_ms(args.cpu_sample, cores)
        scale = N_total / args.cpu_sample
        cpu = {"value": ms_s * scale, "unit": "ms", "cores": cores, "kind": "port",
               "sample": f"one oracle IEKS iteration (NumPy port of the reference, JAX association order, {cores} "
                         f"threads) at N={args.cpu_sample}: {ms_s:.0f} ms, linearly extrapolated x{scale:g}"}

    line = {
        "metric": METRIC, "value": ms_iter, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_iter, "higher_is_better": False, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": dict(workload_config(args), chunk_len=int(L), finite=finite),
        "time_steps_per_s": N_total / (ms_iter * 1e-3),
        "e2e": {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches * args.steps,
        "roofline": roofline, "roofline_hbm": roofline_hbm, "roofline_smooth": roofline_smooth,
        "roofline_iteration": roofline_iter,
        "cpu_baseline": cpu, "clocks": clk,
        "scalars_last": {k: float(v) for k, v in (last.items() if isinstance(last, dict) else
                                                   zip(["nll", "obj", "ssq", "ssq_proper", "not_close"],
                                                       scalars.cpu().tolist()[:5]))},
    }
    emit(line)
    leave()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_native(a)
