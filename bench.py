#!/usr/bin/env python
"""Benchmark of the IEKS hot path: milliseconds per IEKS iteration (fp64), FitzHugh-Nagumo, order 3.

    python bench.py --gpus N --steps K --warmup W               # this repo's CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference   --gpus N --steps K ...   # the reference algorithm on the host cores (oracle port)
    python bench.py --impl gpu_library --steps K ...            # the reference algorithm through stock torch-CUDA library
                                                                # calls on the same GPU (stand-in for its single-GPU JAX time)

Workload (BASELINE.json configs[1] / configs[3]): one IEKS iteration = linearise at the previous trajectory + the
parallel-in-time square-root filter/smoother pass + calibration + convergence reductions (reference
pof/step.py:33-45), on ts = linspace(0, 100, N_total):
    --scaling weak   (default)  N = 2^log2n time points PER GPU (log2n = 20: the configuration the metric is quoted on)
    --scaling strong            N_total = 2^log2n in total (default 22: BASELINE config 4), split over the GPUs
    --log2n E                   BASELINE config 2 (the N sweep) one point at a time
    --start constant|converged  timed iterations continue the IEKS loop from the constant initial trajectory (the
                                worst-conditioned linearisation) or from the converged trajectory
With N GPUs the time axis is sharded contiguously and the per-shard carry elements are exchanged with all-gathers
(pof/sharded.py).  A step touches ~3 GB of HBM per GPU at 2^20 points, far more than the 126 MB L2; for smaller N an
L2-sized buffer is written between timed iterations (see config.l2).  Prints ONE JSON line (rank 0).
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "parallel-in-time-ode-filters_b200")
sys.path[:0] = [ROOT, PKG]

import numpy as np  # noqa: E402

METRIC = "ms per IEKS iteration (fp64)"
D_, d_, q_ = 8, 2, 3
# algorithmic FLOPs per time step per iteration of the REFERENCE formulas at (D, d) = (8, 2)
FLOP_STEP = 68.5e3
# algorithmic HBM bytes per step.  scan: read the compact linearisation [J_f | c] (6 doubles), write the
# step's backward kernel (g: 8, E: 64, untriangularised noise factor: D x (D-d) = 48 doubles); smoother: read the
# backward kernel (120) and the previous mean (8), write the new mean (8) and the calibrated Cholesky factor (64)
BYTES_STEP = {"fold": 6 * 8.0, "scan": (6 + 8 + 64 + 48) * 8.0, "smooth": (120 + 8 + 8 + 64) * 8.0}
L2_BYTES = 126 * 2 ** 20


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference", "gpu_library"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--log2n", type=int, default=None, help="weak: time points per GPU; strong: in total")
    ap.add_argument("--n-time", type=int, default=None, help="(alias) time points per GPU, weak scaling")
    ap.add_argument("--start", default="constant", choices=["constant", "converged"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--e2e-solve", action="store_true", help="also time a full solve() host -> host at N = 2^19")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"],
                    help="f32: the optional fp32 mode (single GPU), reported separately; no parity gate applies")
    a = ap.parse_args()
    if a.log2n is None:
        a.log2n = int(np.log2(a.n_time)) if a.n_time else (22 if a.scaling == "strong" else 20)
    return a


def n_total_of(args, world):
    return 2 ** args.log2n if args.scaling == "strong" else (2 ** args.log2n) * world


def workload_config(args, world, n_total):
    return {
        "workload": f"FitzHugh-Nagumo, IWP order 3 (D=8, d=2), one IEKS iteration, N_total=2^{np.log2(n_total):g} time "
                    f"points on {world} GPU(s) ({args.scaling} scaling), ts=linspace(0,100,N_total), IEKS loop continued "
                    f"from the {args.start} trajectory",
        "n_time_per_gpu": n_total // world, "n_time_total": n_total, "order": q_, "state_dim": D_,
        "start": args.start,
        "sharding": "contiguous time shards, carry all-gather" if world > 1 else "single GPU",
    }


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
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i",
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


# ------------------------------------------------------------------------------------------------ CPU arms (oracle)
def oracle_iterations(n_total, nthreads, steps, warmup, budget_s):
    """The reference's IEKS iteration (NumPy/LAPACK port, JAX association order, oracle/) at the REAL N on the host
    cores, iterations continued from the constant initial trajectory.  Runs `warmup` + `steps` iterations but stops
    early when the time budget is spent.  -> (ms per timed iteration, timed steps run, warm-up steps run, last output)"""
    from oracle import ivps
    from oracle import pof_oracle as O
    from oracle import threaded as OT

    ivp = ivps.fitzhughnagumo()
    ts = np.linspace(0, 100, n_total)
    setup = O.set_up_solver(ivp, ts, q_)
    st = O.get_initial_trajectory(setup)
    t_start = time.perf_counter()
    w_run = 0
    last = None
    for _ in range(warmup):
        t0 = time.perf_counter()
        out = OT.ieks_step(setup, st, nthreads=nthreads)
        st, last = out[0], out
        w_run += 1
        one = time.perf_counter() - t0
        if (time.perf_counter() - t_start) + 2 * one > budget_s:
            break  # keep room for at least one timed step
    k_run = 0
    t_timed = 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        out = OT.ieks_step(setup, st, nthreads=nthreads)
        st, last = out[0], out
        t_timed += time.perf_counter() - t0
        k_run += 1
        if (time.perf_counter() - t_start) + t_timed / k_run > budget_s:
            break
    return t_timed * 1e3 / k_run, k_run, w_run, last


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path (the reference itself needs JAX, which this
    image does not have: the oracle port, all host cores) on the native arm's config.  N_total above 2^22 (the 4- and
    8-GPU weak-scaling points: > 100 GB of scan elements, > 5 min per iteration) is capped at 2^22 and SAID SO -- no
    number is extrapolated."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = args.gpus
    n_cfg = n_total_of(args, world)
    n_run = min(n_cfg, 2 ** 22)
    cores = os.cpu_count() or 1
    ms, k_run, w_run, _ = oracle_iterations(n_run, cores, args.steps, min(args.warmup, 1), budget_s=150.0)
    cfg = workload_config(args, world, n_cfg)
    cfg["n_time_total_run"] = n_run
    cfg["same_config"] = n_run == n_cfg
    sample = (f"oracle port of the reference IEKS iteration (NumPy/LAPACK, JAX association order), {cores} threads, at "
              f"N_total={n_run} (MEASURED, not extrapolated), {w_run} warm-up + {k_run} timed iterations within the "
              f"time budget")
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": k_run,
        "steps_requested": args.steps, "warmup": w_run, "ms_per_step": ms, "higher_is_better": False,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ GPU library arm
def run_gpu_library(args):
    """`--impl gpu_library`: the reference's algorithm through stock torch-CUDA library calls (batched linalg.qr,
    solve_triangular, matmul; recursive odd/even associative scan) on ONE B200 -- oracle/torch_baseline.py.  The stand-in
    for "the reference's single-GPU JAX time" (JAX is not installable here).  None of this repo's kernels run."""
    import torch

    from oracle import ivps
    from oracle import pof_oracle as O
    from oracle import torch_baseline as TB

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", 0)
    n_total = n_total_of(args, 1)
    ts = np.linspace(0, 100, n_total)
    s = O.set_up_solver(ivps.fitzhughnagumo(), ts, q_)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=dev)
    E0, E1, F, QL = t(s["E0"]), t(s["E1"]), t(s["dtm"].F), t(s["dtm"].QL)
    x0m, x0c = t(s["x0"].mean), t(s["x0"].chol)
    row = O.get_initial_trajectory(dict(s, ts=ts[:1])).mean
    means = t(row).expand(n_total, -1).contiguous()
    fj = TB.fhn_f_and_jac()

    def step():
        nonlocal means
        means, chols, nll, obj, ssq, ssqp = TB.ieks_step(fj, E0, E1, F, QL, x0m, x0c, means, calibrate=True)
        return nll

    k_w = min(args.warmup, 2)
    for _ in range(k_w):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = min(args.steps, 5)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    wall = (time.perf_counter() - t0) * 1e3 / steps
    line = {
        "impl": "gpu_library", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": 1, "steps": steps,
        "steps_requested": args.steps, "warmup": k_w, "ms_per_step": ms, "wall_ms_per_step": wall,
        "higher_is_better": False, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": dict(workload_config(args, 1, n_total),
                       library="torch %s: batched linalg.qr / solve_triangular / matmul, recursive odd/even scan "
                               "(oracle/torch_baseline.py)" % torch.__version__),
        "finite": bool(torch.isfinite(means).all().item()),
        "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "gpu_launches": None,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ native arm
def _file_hash(paths):
    h = hashlib.sha256()
    for p in paths:
        h.update(open(p, "rb").read())
    return h.hexdigest()[:16]


def kernel_source_hash():
    """hash of the sources that define the leaf kernels (ties a committed ncu capture to the kernels that ran)"""
    cs = os.path.join(PKG, "csrc")
    return _file_hash([os.path.join(cs, f) for f in ("pof_lane2.cuh", "pof_lane2_kernels.cuh", "pof_small.cuh")])


def ncu_counters():
    """hardware counters of the leaf kernels from the committed `ncu --set full` capture (profiles/r02_ncu_kernels.json)
    -- reported ONLY if the capture was taken from kernels with the same source hash as the ones that just ran"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_kernels.json")))
    except Exception:
        return None
    if t.get("kernel_source_hash") != kernel_source_hash():
        return {"stale": True, "capture_hash": t.get("kernel_source_hash"), "built_hash": kernel_source_hash()}
    return t


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
    affinity = "unchanged"
    if world > 1:
        # one process per GPU: run on (and first-touch the page-locked copy buffers from) the CPUs next to this GPU,
        # so that the end-to-end copies of the ranks do not all go through one socket's memory
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(local).uuid)).encode())
            pynvml.nvmlDeviceSetCpuAffinity(h)
            affinity = "gpu-local (%d cpus)" % len(os.sched_getaffinity(0))
        except Exception as e:  # not fatal: the measurement is still valid, only the copies may be slower
            affinity = "unchanged (%s)" % type(e).__name__
        dist.init_process_group("nccl", device_id=dev)
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)

    import pof.ivp
    from pof import _native as nat
    from pof.convenience import set_up_solver
    from pof.convergence_criteria import crit_scalars
    from pof.parallel_filtsmooth import GraphedCall, GraphedIteration
    from pof.sharded import ShardedPass, shard_bounds

    N_total = n_total_of(args, world)
    n = N_total - 1
    ivp = pof.ivp.fitzhughnagumo()
    tdt = torch.float64 if args.dtype == "f64" else torch.float32
    if tdt != torch.float64 and world > 1:
        raise SystemExit("--dtype f32: single GPU only")

    def make_problem(n_tot):
        """(setup, constant-init means of this rank's rows, k_lo, k_hi)"""
        # set_up_solver only needs the grid spacing and y0: a 2-point grid with the right dt (host, O(1))
        dt = 100.0 / (n_tot - 1)
        setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=np.array([0.0, dt]), order=q_)
        k_lo, k_hi = shard_bounds(n_tot - 1, rank, world)
        rows = (k_hi - k_lo) + (1 if rank == 0 else 0)
        # constant initial trajectory (reference initialization.py:42-56 then PI @ .)
        y0 = ivp.y0.to(dev)
        row = torch.zeros(D_, dtype=torch.float64, device=dev)
        f0 = ivp.f(None, y0)
        for b in range(d_):
            row[b * (q_ + 1)] = y0[b]
            row[b * (q_ + 1) + 1] = f0[b]
        row = setup["PI"] @ row
        if tdt != torch.float64:
            from pof.utils import MVNSqrt

            setup = dict(setup, x0=MVNSqrt(setup["x0"].mean.to(tdt), setup["x0"].chol.to(tdt)))
        return setup, row.repeat(rows, 1).to(tdt).contiguous(), k_lo, k_hi

    exch = {"kind": "none (single GPU)"}

    def make_step(setup, means, chols, k_lo, k_hi, n_tot, calibrate=True):
        """-> (step callable returning the 5 scalars as a device tensor, graph object, chunk_len, launches, ctx)"""
        lin = setup["om"].f._pof_lin
        x0, qL = setup["x0"], setup["_qL"]
        n_loc = k_hi - k_lo
        if world == 1:
            scalars = torch.zeros(nat.NSCALARS, dtype=tdt, device=dev)
            fused = GraphedIteration(x0, qL, lin, means, chols, scalars, calibrate=calibrate)
            L = fused.ws.chunk_len
            launches = 1 + int(nat.LIB.pof_launches_per_pass(n_tot, d_, q_, L, nat.flags()))

            def step():
                fused()
                return scalars

            return step, fused, L, launches, fused.ws.ctx
        sp = ShardedPass(n_tot, d_, q_, qL, rank=rank, world=world, device=dev)
        sp.backend.set_compact(lin["scale0"], lin["scale1"])
        Jc = torch.empty((n_loc, d_ * d_ + d_), dtype=torch.float64, device=dev)
        t1row = 1 if rank == 0 else 0  # local row of the first linearisation point (state k_lo + 1)
        ivp_id, params = lin["builtin"]
        ph, pp = nat.host_doubles(list(params) + [0.0])
        out5 = torch.zeros(nat.NSCALARS, dtype=torch.float64, device=dev)

        def eager():
            nat.check(nat.LIB.pof_linearize_ivp_compact_f64(nat.stream_ptr(), ivp_id, pp, len(params), n_loc, d_, q_,
                                                            lin["scale0"], nat.ptr(means[t1row:]), nat.ptr(Jc)),
                      "linearize")
            res = sp.run(x0.mean, x0.chol, Jc, None, means, chols, calibrate=calibrate)
            if "scalars" in res:
                return res["scalars"]
            out5[:5].copy_(torch.stack([res["nll"], res["obj"], res["ssq"], res["ssq_proper"], res["not_close"]]))
            return out5

        fused = GraphedCall(eager)
        exch["kind"] = sp.exchange
        L = sp.backend.chunk_len
        # linearise + the kernels of a pass + the separate up-sweep launch of stage A + the three exchange kernels
        launches = 1 + int(nat.LIB.pof_launches_per_pass(n_loc + 1, d_, q_, L, nat.flags())) + 1 + 3
        return (lambda: fused()), fused, L, launches, sp.backend.ws.ctx

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # inputs smaller than L2 would otherwise be timed out of cache: write an L2-sized buffer between timed iterations
    per_gpu_bytes = (N_total // world) * 2700
    flush = None
    if per_gpu_bytes < 4 * L2_BYTES:
        flush = torch.empty(2 * L2_BYTES // 8, dtype=torch.float64, device=dev)

    def timed(fn, steps):
        barrier()
        if flush is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        else:
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            for a, b in evs:
                flush.fill_(1.0)
                a.record()
                fn()
                b.record()
            barrier()
            ms = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    setup, means0, k_lo, k_hi = make_problem(N_total)
    n_loc = k_hi - k_lo
    rows = means0.shape[0]
    means = means0.clone()
    chols = torch.empty((rows, D_, D_), dtype=tdt, device=dev)
    step, fused, L, launches, ctx = make_step(setup, means, chols, k_lo, k_hi, N_total)

    its_to_converge = None
    if args.start == "converged":  # run the IEKS loop to its stopping rule first (reference solver.py:36-45)
        obj_old = nll_old = 0.0
        for k in range(1000):
            sc = step().cpu()
            nll, obj, bad = float(sc[0]), float(sc[1]), float(sc[4])
            if k >= 1 and crit_scalars(obj, obj_old, nll, nll_old, bad):
                break
            nll_old, obj_old = nll, obj
        its_to_converge = k + 1

    # ---- device-resident timing (value); per-segment CUDA events in a separate eager loop
    for _ in range(args.warmup):
        step()
    ctx.profile_enable(True)
    ms_eager = timed(step, max(3, args.steps // 4))
    seg_raw = ctx.profile_read()
    ctx.profile_enable(False)
    barrier()
    fused.capture()
    step()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms_iter = timed(step, args.steps)
    clk = clocks.stop() if rank == 0 else None
    last = step().cpu().numpy()
    torch.cuda.synchronize()
    finite = bool(torch.isfinite(means).all().item())

    # ---- end to end through the public call with HOST buffers: H2D of the previous trajectory means from pinned
    # memory, the iteration, D2H of the projected solution means E0 m (N,d) and the scalars
    h_means = torch.empty((rows, D_), dtype=tdt).pin_memory()
    h_means.copy_(means.cpu())
    h_y = torch.empty((rows, d_), dtype=tdt).pin_memory()
    h_sc = torch.empty(nat.NSCALARS, dtype=tdt).pin_memory()
    ymean = torch.empty((rows, d_), dtype=tdt, device=dev)
    esz = 8 if tdt == torch.float64 else 4

    def e2e_step():
        means.copy_(h_means, non_blocking=True)
        sc = step()
        nat.check(nat.fn("pof_project", tdt)(nat.stream_ptr(), rows, d_, q_, setup["_scale0"], None, nat.ptr(means),
                                             None, nat.ptr(ymean), None), "project")
        h_y.copy_(ymean, non_blocking=True)
        h_sc.copy_(sc, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    for _ in range(2):
        e2e_step()
    ms_e2e_serial = timed(e2e_step, max(3, args.steps // 2))
    h2d = rows * D_ * esz
    d2h = rows * d_ * esz + nat.NSCALARS * esz

    # The same end-to-end steps DOUBLE-BUFFERED: every step still copies its inputs host -> device and its results device
    # -> host, but the copies of step i+1 / i-1 run on two copy streams while the kernels of step i execute (two staging
    # buffers each way, events for reuse).  This is how a host-resident caller streams independent steps.
    main = torch.cuda.current_stream()
    s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
    stage = [torch.empty_like(means) for _ in range(2)]
    ym = [torch.empty_like(ymean) for _ in range(2)]
    scs = [torch.empty(nat.NSCALARS, dtype=tdt, device=dev) for _ in range(2)]
    hy = [torch.empty((rows, d_), dtype=tdt).pin_memory() for _ in range(2)]
    hs = [torch.empty(nat.NSCALARS, dtype=tdt).pin_memory() for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_used = [torch.cuda.Event() for _ in range(2)]
    ev_out = [torch.cuda.Event() for _ in range(2)]
    ev_read = [torch.cuda.Event() for _ in range(2)]
    pipe_i = [0]

    def e2e_pipe_step():
        i = pipe_i[0]
        b = i & 1
        pipe_i[0] += 1
        with torch.cuda.stream(s_h2d):
            if i >= 2:
                s_h2d.wait_event(ev_used[b])  # the staging buffer has been consumed by step i-2
            stage[b].copy_(h_means, non_blocking=True)
            ev_in[b].record(s_h2d)
        main.wait_event(ev_in[b])
        means.copy_(stage[b], non_blocking=True)
        ev_used[b].record(main)
        sc = step()
        if i >= 2:
            main.wait_event(ev_read[b])  # step i-2's results have left the device buffers
        nat.check(nat.fn("pof_project", tdt)(nat.stream_ptr(), rows, d_, q_, setup["_scale0"], None, nat.ptr(means),
                                             None, nat.ptr(ym[b]), None), "project")
        scs[b].copy_(sc, non_blocking=True)
        ev_out[b].record(main)
        with torch.cuda.stream(s_d2h):
            s_d2h.wait_event(ev_out[b])
            hy[b].copy_(ym[b], non_blocking=True)
            hs[b].copy_(scs[b], non_blocking=True)
            ev_read[b].record(s_d2h)

    def e2e_pipe_timed(steps):
        barrier()
        pipe_i[0] = 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            e2e_pipe_step()
        main.wait_stream(s_d2h)  # the last results are on the host before the clock stops
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()) / steps

    e2e_pipe_timed(4)
    ms_e2e = e2e_pipe_timed(max(6, args.steps))

    # ---- parity of the path that was just timed, against the CPU oracle, at a size the oracle finishes in seconds
    # (N_total = 2^15, same sharding, ONE uncalibrated pass from the constant trajectory); max over ranks
    parity = None
    if not args.no_parity and tdt == torch.float64:
        parity = parity_check(torch, dist, nat, make_problem, make_step, rank, world, dev)

    # ---- FP64 peak of this device (no FP64 figure in MEASURED_PEAKS.json)
    tf = np.zeros(1)
    nat.check(nat.LIB.pof_measure_dfma_tflops(nat.stream_ptr(), tf.ctypes.data_as(nat._c_dp)), "dfma peak")
    fp64_peak = float(tf[0])

    # ---- a full solve() from host inputs to host outputs (optional: --e2e-solve), N = 2^19 like the published V100 run
    e2e_solve = None
    if args.e2e_solve and world == 1:
        from pof.solver import solve

        ts = np.linspace(0, 100, 2 ** 19)
        outs = {}
        # result buffers in page-locked host memory, allocated once by the caller (as a service that solves repeatedly
        # would); the copies into them are inside the timed region
        hm = torch.empty((2 ** 19, d_), dtype=torch.float64, pin_memory=True)
        hc = torch.empty((2 ** 19, d_, D_), dtype=torch.float64, pin_memory=True)
        for rep in range(4):  # first run pays one-off costs (allocations, graph capture); then the best of three
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ys, info = solve(f=ivp.f, y0=ivp.y0, ts=ts, order=q_, init="constant", maxiters=1000)
            hm.copy_(ys.mean, non_blocking=True)
            hc.copy_(ys.chol, non_blocking=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if rep <= 1 or dt < outs[1][0]:
                outs[min(rep, 1)] = (dt, info["iterations"], hm.shape, hc.shape)
        e2e_solve = {"value": outs[1][0], "unit": "s", "first_call_s": outs[0][0], "iterations": outs[1][1],
                     "n_time": 2 ** 19, "outputs": "means (N,d) + Cholesky factors (N,d,D) copied to page-locked host memory; best of 3",
                     "d2h_bytes": int(2 ** 19 * (d_ + d_ * D_) * 8),
                     "published_reference": "54.74 s, 112 iterations, V100 JAX (BASELINE.md)"}

    def teardown():
        """release the captured NCCL work, then destroy the process group; a teardown that does not return within 20 s
        (seen once with graph-captured collectives) must not hang the job: the result line is out by then"""
        if world <= 1:
            return
        fused.graph = None
        torch.cuda.synchronize()
        th = threading.Thread(target=dist.destroy_process_group, daemon=True)
        th.start()
        th.join(20.0)
        if th.is_alive():
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(0)

    if rank != 0:
        teardown()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (
        6650.0, "fallback (B200_PROFILING.md)")
    seg = {nm: (v[0] / max(1, v[1])) for nm, v in seg_raw.items()}
    ncu = ncu_counters()
    dominant = max(("fold", "scan", "smooth"), key=lambda k: seg[k])
    kname = {"fold": "k_lane2_fold<2,3>", "scan": "k_lane2_scan<2,3>", "smooth": "k_lane2_smooth<2,3,0>"}

    def hbm_view(k):
        t = seg[k]
        ach = BYTES_STEP[k] * n_loc / (t * 1e-3) / 1e9 if t > 0 else None
        cap = (ncu or {}).get(kname[k], {}) if ncu and not ncu.get("stale") else {}
        return {"kernel": kname[k], "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                "frac": ach / hbm_peak if ach else None, "traffic": cap.get("dram_bytes"), "peak_source": hbm_src,
                "algorithmic_bytes_per_step": BYTES_STEP[k], "avg_launch_ms": t,
                "share_of_step": t / ms_eager if ms_eager else None}

    # The dominant kernels are FP64-pipe / latency bound, not HBM bound: the executed-work fraction of
    # a kernel is its FP64-pipe utilisation, a hardware counter -- reported from the committed ncu capture only if that
    # capture was taken from the same kernel sources.  `roofline` = the dominant kernel against the roof that binds the
    # PATH (fp64); frac = executed FP64 work / measured DFMA peak; its HBM view sits beside it in `roofline_hbm`.
    cap = (ncu or {}).get(kname[dominant], {}) if ncu and not ncu.get("stale") else {}
    fp64_frac = cap.get("fp64_pipe_pct", None)
    roofline = {
        "kernel": kname[dominant], "bound": "fp64",
        "achieved": fp64_frac / 100.0 * fp64_peak if fp64_frac is not None else None, "peak": fp64_peak,
        "unit": "TFLOP/s", "frac": fp64_frac / 100.0 if fp64_frac is not None else None,
        "traffic": cap.get("dram_bytes"), "avg_launch_ms": seg[dominant],
        "share_of_step": seg[dominant] / ms_eager if ms_eager else None,
        "peak_source": "DFMA loop measured in this run (pof_measure_dfma_tflops; nominal 37 TFLOP/s)",
        "frac_source": ("executed-work view: sm__inst_executed_pipe_fp64 of the committed ncu --set full capture "
                        "(profiles/r02_ncu_kernels.json), same kernel sources (hash %s)" % kernel_source_hash())
        if fp64_frac is not None else "no ncu capture of these kernel sources committed: executed-work fraction unknown",
    }
    iter_tflops = FLOP_STEP * n_loc / (ms_iter * 1e-3) / 1e12
    roofline_iter = {
        "scope": "whole IEKS iteration, USEFUL work: reference-formula FLOPs of SURVEY 8d (68.5 kFLOP per step) over "
                 "the measured time; exceeds the executed-work fraction because a leaf recursion reaches the result of a "
                 "general combine with ~8x fewer flops",
        "bound": "fp64", "achieved": iter_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": iter_tflops / fp64_peak,
        "algorithmic_flop_per_step": FLOP_STEP, "segments_ms": seg, "ms_per_step_eager_launches": ms_eager,
        "hbm_bytes_per_step_algorithmic": sum(BYTES_STEP.values()),
        "hbm_frac_of_peak": sum(BYTES_STEP.values()) * n_loc / (ms_iter * 1e-3) / 1e9 / hbm_peak,
    }

    cpu = None
    if world == 1 and not args.no_cpu_baseline and args.dtype == "f64":
        cores = os.cpu_count() or 1
        n_cpu = min(N_total, 2 ** 20)
        ms_c, k_run, w_run, _ = oracle_iterations(n_cpu, cores, 1, 0, budget_s=120.0)
        cpu = {"value": ms_c, "unit": "ms", "cores": cores, "kind": "port",
               "sample": f"one oracle IEKS iteration (NumPy port of the reference, JAX association order, {cores} "
                         f"threads) at N={n_cpu}, MEASURED at that N (no extrapolation)", "n_time": n_cpu}

    cfg = dict(workload_config(args, world, N_total), chunk_len=int(L), finite=finite, exchange=exch["kind"],
               cpu_affinity=affinity, l2="working set per step (~3 GB at 2^20 points) exceeds L2 (126 MB); no explicit flush" if flush is None
               else "L2 flushed between timed iterations (a 252 MB buffer is written before each)")
    if its_to_converge is not None:
        cfg["iterations_to_converge_before_timing"] = its_to_converge
    line = {
        "metric": METRIC if args.dtype == "f64" else "ms per IEKS iteration (fp32: optional mode, reported separately)",
        "value": ms_iter, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_iter, "higher_is_better": False, "scaling": args.scaling, "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic", "config": cfg,
        "time_steps_per_s": N_total / (ms_iter * 1e-3),
        "e2e": {"value": ms_e2e, "unit": "ms", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "mode": "double-buffered: every step copies its inputs H2D (pinned) and its results D2H; the copies of "
                        "neighbouring steps overlap this step's kernels on two copy streams",
                "serial_value": ms_e2e_serial,
                "serial_mode": "H2D, kernels, D2H strictly one after the other, host sync per step"},
        "gpu_launches": launches * args.steps, "gpu_launches_per_step": launches,
        "roofline": roofline, "roofline_hbm": hbm_view(dominant), "roofline_smooth": hbm_view("smooth"),
        "roofline_iteration": roofline_iter,
        "parity": parity, "cpu_baseline": cpu, "e2e_solve": e2e_solve, "clocks": clk,
        "scalars_last": dict(zip(["nll", "obj", "ssq", "ssq_proper", "not_close"], [float(v) for v in last[:5]])),
    }
    emit(line)
    teardown()


def parity_check(torch, dist, nat, make_problem, make_step, rank, world, dev):
    """one uncalibrated pass at N_total = 2^15 through the SAME code path (graph-free), every rank against the oracle
    rows it owns; gates of SURVEY 8c: outputs 1e-9, projected covariance 1e-7, sigma^2 (sign-invariant) 1e-8 ... 1e-6"""
    from oracle import ivps
    from oracle import pof_oracle as O
    from oracle import threaded as OT

    n_par = 2 ** 15
    setup, means, k_lo, k_hi = make_problem(n_par)
    rows = means.shape[0]
    chols = torch.empty((rows, D_, D_), dtype=torch.float64, device=dev)
    step, fused, _, _, _ = make_step(setup, means, chols, k_lo, k_hi, n_par, calibrate=False)
    sc = step().cpu().numpy()
    torch.cuda.synchronize()
    ts = np.linspace(0, 100, n_par)
    os_ = O.set_up_solver(ivps.fitzhughnagumo(), ts, q_)
    ost = O.get_initial_trajectory(os_)
    oout, onll, oobj, ossq, ossqp = OT.ieks_step(os_, ost, calibrate=False,
                                                  nthreads=max(1, (os.cpu_count() or 1) // world))
    r0 = 0 if rank == 0 else k_lo + 1
    sl = slice(r0, k_hi + 1)
    E0 = os_["E0"]
    y, yo = means.cpu().numpy() @ E0.T, oout.mean[sl] @ E0.T
    Lg, Lo = chols.cpu().numpy(), oout.chol[sl]
    Pg = E0 @ (Lg @ np.swapaxes(Lg, -1, -2)) @ E0.T
    Po = E0 @ (Lo @ np.swapaxes(Lo, -1, -2)) @ E0.T
    scale = np.abs(oout.mean @ E0.T).max(axis=0)
    Pall = E0 @ (oout.chol @ np.swapaxes(oout.chol, -1, -2)) @ E0.T
    errs = torch.tensor([np.max(np.abs(y - yo) / scale), np.abs(Pg - Po).max() / np.abs(Pall).max(),
                         abs(sc[3] - ossqp) / abs(ossqp), abs(sc[0] - onll) / abs(onll), abs(sc[1] - oobj) / abs(oobj)],
                        dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    e = [float(v) for v in errs.cpu()]
    ok = e[0] <= 1e-9 and e[1] <= 1e-7 and e[2] <= 1e-6 and e[3] <= 1e-9 and e[4] <= 1e-9
    return {"n_time_total": n_par, "max_rel_y": e[0], "max_rel_cov": e[1], "ssq_proper_rel": e[2], "nll_rel": e[3],
            "obj_rel": e[4], "ok": bool(ok),
            "note": "this run's code path (same sharding / exchanges) vs the CPU oracle, one pass from the constant "
                    "trajectory; max over ranks"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "gpu_library":
        run_gpu_library(a)
    else:
        run_native(a)
