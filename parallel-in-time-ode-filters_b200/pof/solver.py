"""Solver API (reference pof/solver.py:11-96): `solve` and `sequential_eks_solve`, keyword-only, same defaults,
same `(MVNSqrt(mean (N,d), chol (N,d,D)), info_dict)` results -- as torch CUDA tensors."""
import ctypes

import torch

from . import _native as nat
from .convenience import get_initial_trajectory, set_up_solver
from .convergence_criteria import crit_scalars
from .parallel_filtsmooth import GraphedIteration, run_pass
from .step import linearize_into
from .utils import MVNSqrt


def solve(*, f, y0, ts, order, init="prior", calibrate=True, maxiters=10_000, sequential=False, chunk_len=None,
          dtype=torch.float64):
    """reference solver.py:11-73.  The IEKS loop keeps every array on the device; per iteration the host reads back
    five scalars (nll, obj, sigma^2, #means not close, NaN state) to evaluate the reference's stopping rule.

    dtype=torch.float32 selects the optional fp32 mode (the reference without JAX_ENABLE_X64): the set-up stays fp64,
    the iterations run the fp32 build of the register-resident kernels (built-in `pof.ivp` problems with D <= 16);
    results are float32 tensors.  The mean-convergence rule keeps the reference's absolute tolerance 1e-8, which fp32
    cannot meet: the loop then ends on the objective rule (rtol 1e-6) or `maxiters`."""
    setup = set_up_solver(f=f, y0=y0, ts=ts, order=order)
    x0, om, dev = setup["x0"], setup["om"], setup["_device"]
    lin = om.f._pof_lin
    d, q = lin["d"], order
    states = get_initial_trajectory(setup, method=init, means_only=True)
    if dtype != torch.float64:
        if lin["builtin"] is None or sequential:
            raise NotImplementedError("fp32 mode: built-in pof.ivp vector fields, parallel pass only")
        x0 = MVNSqrt(x0.mean.to(dtype), x0.chol.to(dtype))

    means = states.mean.to(dtype).contiguous()
    N, D = means.shape
    n = N - 1
    chols = torch.empty((N, D, D), dtype=dtype, device=dev)
    if lin["builtin"] is None:
        H = torch.empty((n, d, D), dtype=torch.float64, device=dev)
        c = torch.empty((n, d), dtype=torch.float64, device=dev)
    scalars = torch.zeros(nat.NSCALARS, dtype=dtype, device=dev)
    if sequential:
        chunk_len = n
    elif chunk_len is None:
        chunk_len = nat.default_chunk_len(N, d, q, dev.index)

    nll = obj = ssq = 0.0
    nll_old = obj_old = 0.0
    k = 0
    fused = None
    # Built-in vector field on the register-resident kernels: the whole loop -- body AND stopping rule -- runs on the
    # device (the reference's lax.while_loop, solver.py:36-57).  After one eager iteration the remaining ones are ONE
    # launch of a CUDA graph whose body sits in a WHILE conditional node (`pof_ieks_loop_create`); the host reads the
    # loop state once, at the end.  Where the driver cannot build conditional nodes, a captured single iteration
    # (`pof_ieks_loop_step`) is replayed `burst` times between two reads of the loop state instead: iterations
    # enqueued after the rule has fired are no-ops, so results and iteration counts are the same either way.
    D_ = d * (q + 1)
    device_loop = (lin["builtin"] is not None and not sequential
                   and bool(nat.LIB.pof_shard_exchange_supported(D_, nat.flags())))
    if device_loop:
        loop_state = torch.zeros(8, dtype=dtype, device=dev)
        fused = GraphedIteration(x0, setup["_qL"], lin, means, chols, scalars, calibrate=True, chunk_len=chunk_len,
                                 loop_state=loop_state, maxiters=maxiters)
        read_state = lambda: torch.cat([loop_state[:4], scalars]).cpu()  # synchronises
        fused()
        if nat.USE_LOOP_GRAPH and fused.capture_loop():
            fused.launch_loop()
            st = read_state()
        else:
            burst = 1 if N >= 2 ** 17 else 8
            fused.capture()
            st = read_state()
            while float(st[0]) == 0.0:
                for _ in range(burst):
                    fused()
                st = read_state()
        k = int(st[1])
        sc = st[4:]
        nll, obj, ssq = float(sc[nat.S_NLL]), float(sc[nat.S_OBJ]), float(sc[nat.S_SSQ])
    while not device_loop:
        if k >= 1:
            converged = crit_scalars(obj, obj_old, nll, nll_old, n_bad)
            if converged or not (k <= maxiters):
                break
        nll_old, obj_old = nll, obj
        # body: ieks_step (solver.py:48-55); it always calibrates inside the loop (calibrate is not forwarded)
        if lin["builtin"] is not None:  # fused f / Jacobian + pass, H never materialised; graph-replayed
            if fused is None:
                fused = GraphedIteration(x0, setup["_qL"], lin, means, chols, scalars, calibrate=True,
                                         chunk_len=chunk_len)
            elif fused.graph is None and k >= 1:
                fused.capture()  # note: capturing runs no kernels; the replay below is iteration k+1
            fused()
        else:  # user f: autodiff linearisation on the device, then the same pass
            linearize_into(lin, means, H, c)
            run_pass(x0, setup["_qL"], H, c, means, chols, d=d, q=q, calibrate=True, chunk_len=chunk_len,
                     scalars=scalars)
        sc = scalars.cpu()
        nll, obj, ssq, n_bad = float(sc[nat.S_NLL]), float(sc[nat.S_OBJ]), float(sc[nat.S_SSQ]), float(
            sc[nat.S_NOT_CLOSE])
        k += 1

    if sequential:
        # the reference's sequential pass returns ell = +sum log-likelihood where the parallel one returns the
        # negative sum (quirk Q7: sequential_filtsmooth/filter.py:91 vs parallel_filtsmooth/filter.py:101)
        nll = -nll
    info_dict = {
        "iterations": k, "nll": nll, "obj": obj, "sigma_squared": ssq, "calibrated": False,
        "sigma_squared_proper": float(sc[nat.S_SSQ_PROPER]),
    }
    if calibrate:
        info_dict["calibrated"] = True
    # final calibration (the second one, solver.py:66-69) fused with the E0 projection (solver.py:71)
    ymean = torch.empty((N, d), dtype=dtype, device=dev)
    ychol = torch.empty((N, d, D), dtype=dtype, device=dev)
    mult = scalars[nat.S_CSCALE:nat.S_CSCALE + 1] if calibrate else None
    rc = nat.fn("pof_project", dtype)(nat.stream_ptr(), N, d, q, setup["_scale0"], nat.ptr(mult), nat.ptr(means),
                                 nat.ptr(chols), nat.ptr(ymean), nat.ptr(ychol))
    nat.check(rc, "pof_project_f64")
    return MVNSqrt(ymean, ychol), info_dict


def sequential_eks_solve(*, f, y0, ts, order, return_full_states=False, calibrate=True):
    """reference solver.py:76-96: one extended Kalman filter pass relinearised at the predicted mean, then RTS."""
    from .sequential_filtsmooth.eks import eks_filtsmooth

    setup = set_up_solver(f=f, y0=y0, ts=ts, order=order)
    states, nll, obj, ssq = eks_filtsmooth(setup)
    info_dict = {"nll": nll, "obj": obj, "sigma_squared": ssq, "calibrated": False}
    if calibrate:
        states = MVNSqrt(states.mean, float(ssq) ** 0.5 * states.chol)
        info_dict["calibrated"] = True
    M = setup["P"] if return_full_states else setup["E0"]
    return MVNSqrt(states.mean @ M.T, torch.einsum("ij,njk->nik", M, states.chol)), info_dict
