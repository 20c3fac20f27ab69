"""A batch of independent IVPs solved concurrently (the shape of the reference's benchmark runner,
experiments/3_work_precision_diagram/run_benchmark.py:114-225, which solves its problems one after the other).

Many independent IVPs are a second, trivially data-parallel axis: below N ~ 2^16 one IEKS iteration is bound by the
latency of its tree sweeps and uses a fraction of the GPU, so B problems are advanced in lockstep -- every problem on
its own CUDA stream with its own workspace, execution context and CUDA graph (`GraphedIteration`), ONE host
synchronisation per round for the stopping rules of all of them.  Each problem runs exactly the iterations
`pof.solver.solve` would run for it (same kernels, same stopping rule), so results are identical to separate calls.
"""
import torch

from . import _native as nat
from .convenience import get_initial_trajectory, set_up_solver
from .convergence_criteria import crit_scalars
from .parallel_filtsmooth import GraphedIteration
from .utils import MVNSqrt


def solve_batch(problems, *, order, init="prior", calibrate=True, maxiters=10_000):
    """problems: sequence of dicts with keys f, y0, ts (built-in `pof.ivp` vector fields).
    -> list of (MVNSqrt(mean (N,d), chol (N,d,D)), info_dict), one per problem, as `pof.solver.solve` returns them."""
    jobs = []
    for p in problems:
        setup = set_up_solver(f=p["f"], y0=p["y0"], ts=p["ts"], order=order)
        lin = setup["om"].f._pof_lin
        if lin["builtin"] is None:
            raise NotImplementedError("solve_batch: built-in pof.ivp vector fields (the fused, graph-replayed iteration)")
        dev = setup["_device"]
        means = get_initial_trajectory(setup, method=init, means_only=True).mean.contiguous()
        N, D = means.shape
        chols = torch.empty((N, D, D), dtype=torch.float64, device=dev)
        scalars = torch.zeros(nat.NSCALARS, dtype=torch.float64, device=dev)
        it = GraphedIteration(setup["x0"], setup["_qL"], lin, means, chols, scalars, calibrate=True)
        jobs.append(dict(setup=setup, lin=lin, means=means, chols=chols, scalars=scalars, it=it,
                         stream=torch.cuda.Stream(device=dev), k=0, nll=0.0, obj=0.0, ssq=0.0, nll_old=0.0,
                         obj_old=0.0, bad=1.0, done=False, sc=None))
    main = torch.cuda.current_stream()
    while True:
        active = []
        for j in jobs:
            if j["done"]:
                continue
            if j["k"] >= 1 and (crit_scalars(j["obj"], j["obj_old"], j["nll"], j["nll_old"], j["bad"])
                                or not (j["k"] <= maxiters)):
                j["done"] = True
                continue
            active.append(j)
        if not active:
            break
        for j in active:
            j["nll_old"], j["obj_old"] = j["nll"], j["obj"]
            j["stream"].wait_stream(main)
            with torch.cuda.stream(j["stream"]):
                if j["it"].graph is None and j["k"] >= 1:
                    j["it"].capture()
                j["it"]()
        for j in active:
            main.wait_stream(j["stream"])
        allsc = torch.stack([j["scalars"] for j in active]).cpu()  # the round's only host synchronisation
        for j, sc in zip(active, allsc):
            j["nll"], j["obj"], j["ssq"] = float(sc[nat.S_NLL]), float(sc[nat.S_OBJ]), float(sc[nat.S_SSQ])
            j["bad"] = float(sc[nat.S_NOT_CLOSE])
            j["sc"] = sc
            j["k"] += 1
    out = []
    for j in jobs:
        setup, lin = j["setup"], j["lin"]
        N, D = j["means"].shape
        d, q = lin["d"], lin["q"]
        dev = setup["_device"]
        ymean = torch.empty((N, d), dtype=torch.float64, device=dev)
        ychol = torch.empty((N, d, D), dtype=torch.float64, device=dev)
        mult = j["scalars"][nat.S_CSCALE:nat.S_CSCALE + 1] if calibrate else None
        nat.check(nat.LIB.pof_project_f64(nat.stream_ptr(), N, d, q, setup["_scale0"], nat.ptr(mult),
                                          nat.ptr(j["means"]), nat.ptr(j["chols"]), nat.ptr(ymean), nat.ptr(ychol)),
                  "pof_project_f64")
        info = {"iterations": j["k"], "nll": j["nll"], "obj": j["obj"], "sigma_squared": j["ssq"],
                "calibrated": bool(calibrate), "sigma_squared_proper": float(j["sc"][nat.S_SSQ_PROPER])}
        out.append((MVNSqrt(ymean, ychol), info))
    return out
