"""Parallel-in-time square-root filter / smoother (reference pof/parallel_filtsmooth/).

Same entry points and return values as the reference; every one of them runs the CUDA pass
(pof_linear_filtsmooth_f64: blocked prefix / suffix scans) -- nothing is computed on the host.
"""
import ctypes

import numpy as np
import torch

from .. import _native as nat
from ..observations import AffineModel
from ..transitions import IWP, TransitionModel, preconditioned_discretize
from ..utils import MVNSqrt


def _model_dims(dtm: TransitionModel, dom: AffineModel):
    """-> (n, d, q, D, qL, dense): dense = None for the preconditioned IWP of pof.transitions (the specialised kernels),
    else contiguous per-step (F (n,D,D), QL (n,D,D)) for the general kernels (`pof_linear_filtsmooth_general_f64`)."""
    n, d, D = dom.H.shape
    if D % d != 0:
        raise ValueError("state dimension must be d*(q+1)")
    q = D // d - 1
    Fi, QLi = preconditioned_discretize(IWP(num_derivatives=q, wiener_process_dimension=d))
    qL = np.ascontiguousarray(QLi[: q + 1, : q + 1])

    def is_const(M, ref):
        M0 = M[0] if M.dim() == 3 else M
        if not np.allclose(M0.detach().cpu().numpy(), ref, rtol=0, atol=1e-12):
            return False
        return M.dim() == 2 or bool((M == M0).all())

    if is_const(dtm.F, Fi) and is_const(dtm.QL, QLi):
        return n, d, q, D, qL, None
    if not bool((torch.triu(dtm.QL, 1) == 0).all()):
        raise ValueError("QL must be lower triangular (a Cholesky factor of the process-noise covariance)")
    rep = lambda M: (M if M.dim() == 3 else M.unsqueeze(0).expand(n, D, D)).contiguous()
    return n, d, q, D, qL, (rep(dtm.F), rep(dtm.QL))


def _noise(dom: AffineModel):
    """cholR as a contiguous (n,d,d) tensor if the observations are noisy, else None (reference observations.py:23-33)"""
    if dom.cholR is None or not bool((dom.cholR != 0).any()):
        return None
    return dom.cholR.contiguous()


def run_pass(x0: MVNSqrt, qL, H, c, means_io, chols, *, d, q, calibrate, chunk_len=None, fmeans=None, fchols=None,
             scalars=None, cholR=None, dense=None, ws=None):
    """One filter+smoother pass on device buffers (the call `solve` makes every iteration).  cholR (n,d,d): noisy
    observations; dense = (F (n,D,D), QL (n,D,D)): general per-step transition models -- both served by the
    large-state kernels (`pof_linear_filtsmooth_general_f64`)."""
    N = means_io.shape[0]
    general = cholR is not None or dense is not None
    Fd, QLd = dense if dense is not None else (None, None)
    nat.require_cuda(x0.mean, x0.chol, H, c, means_io, chols, fmeans, fchols, cholR, Fd, QLd)
    dev = means_io.device
    if chunk_len is None:
        chunk_len = (nat.default_chunk_len_tile if general else nat.default_chunk_len)(N, d, q, dev.index)
    dt = means_io.dtype
    if ws is None:
        ws = nat.Workspace.get(N, d, q, chunk_len, dev, dt)
    elif not ws.matches(N, d, q, chunk_len, dt):
        raise nat.NativeError("workspace was created for a different problem shape")
    if scalars is None:
        scalars = torch.zeros(nat.NSCALARS, dtype=dt, device=dev)
    qLh, qLp = nat.host_doubles(qL)
    if general:
        if dt != torch.float64:
            raise nat.NativeError("noisy observations / general transition models: fp64 only (large-state kernels)")
        rc = nat.LIB.pof_linear_filtsmooth_general_f64(
            nat.stream_ptr(), ws.ctx.ptr, nat.flags(), N, d, q, int(chunk_len), qLp, nat.ptr(Fd), nat.ptr(QLd), nat.ptr(x0.mean),
            nat.ptr(x0.chol), nat.ptr(H), nat.ptr(c), nat.ptr(cholR), nat.ptr(means_io), nat.ptr(chols),
            nat.ptr(fmeans), nat.ptr(fchols), int(bool(calibrate)), nat.ptr(scalars),
            ws.ws_ptr, ws.nbytes)
        nat.check(rc, "pof_linear_filtsmooth_general_f64")
        return scalars
    rc = nat.fn("pof_linear_filtsmooth", dt)(
        nat.stream_ptr(), ws.ctx.ptr, nat.flags(), N, d, q, int(chunk_len), qLp, nat.ptr(x0.mean), nat.ptr(x0.chol), nat.ptr(H), nat.ptr(c),
        nat.ptr(means_io), nat.ptr(chols), nat.ptr(fmeans), nat.ptr(fchols), int(bool(calibrate)), nat.ptr(scalars),
        ws.ws_ptr, ws.nbytes)
    nat.check(rc, "pof_linear_filtsmooth_f64")
    return scalars


def run_iteration(x0: MVNSqrt, qL, lin, means_io, chols, *, calibrate, chunk_len=None, scalars=None, ws=None,
                  loop_state=None, maxiters=10_000):
    """One fused IEKS iteration for a built-in IVP (linearise + pass, `pof_ieks_iteration_f64`): the body of the
    reference's while loop (pof/solver.py:48-55 -> pof/step.py:33-45).  `lin` is `om.f._pof_lin`.
    With `loop_state` (8 zeros on the device) the stopping rule is evaluated on the device too
    (`pof_ieks_loop_step_f64`): once it has fired, further calls are no-ops."""
    N = means_io.shape[0]
    d, q = lin["d"], lin["q"]
    nat.require_cuda(x0.mean, x0.chol, means_io, chols)
    dev = means_io.device
    if chunk_len is None:
        chunk_len = nat.default_chunk_len(N, d, q, dev.index)
    dt = means_io.dtype
    if ws is None:
        ws = nat.Workspace.get(N, d, q, chunk_len, dev, dt)
    elif not ws.matches(N, d, q, chunk_len, dt):
        raise nat.NativeError("workspace was created for a different problem shape")
    if scalars is None:
        scalars = torch.zeros(nat.NSCALARS, dtype=dt, device=dev)
    ivp_id, params = lin["builtin"]
    ph, pp = nat.host_doubles(list(params) + [0.0])
    qLh, qLp = nat.host_doubles(qL)
    if loop_state is not None:
        nat.require_cuda(loop_state, means_io)
        rc = nat.fn("pof_ieks_loop_step", dt)(
            nat.stream_ptr(), ws.ctx.ptr, nat.flags(), ivp_id, pp, len(params), N, d, q, int(chunk_len), qLp,
            lin["scale0"], lin["scale1"], nat.ptr(x0.mean), nat.ptr(x0.chol), nat.ptr(means_io), nat.ptr(chols),
            int(bool(calibrate)), nat.ptr(scalars), nat.ptr(loop_state), int(maxiters), ws.ws_ptr, ws.nbytes)
        nat.check(rc, "pof_ieks_loop_step")
        return scalars
    rc = nat.fn("pof_ieks_iteration", dt)(
        nat.stream_ptr(), ws.ctx.ptr, nat.flags(), ivp_id, pp, len(params), N, d, q, int(chunk_len), qLp, lin["scale0"], lin["scale1"],
        nat.ptr(x0.mean), nat.ptr(x0.chol), nat.ptr(means_io), nat.ptr(chols), int(bool(calibrate)),
        nat.ptr(scalars), ws.ws_ptr, ws.nbytes)
    nat.check(rc, "pof_ieks_iteration_f64")
    return scalars


class GraphedIteration:
    """The fused IEKS iteration captured once into a CUDA graph and replayed (the ~60 kernel launches of a pass
    cost ~0.1 ms of launch gaps otherwise; CUDA streams and graphs replace the reference's jit-compiled loop body).
    All buffers are fixed: `means` is updated in place by every replay, `scalars` holds the iteration's scalars."""

    def __init__(self, x0, qL, lin, means, chols, scalars, *, calibrate=True, chunk_len=None, loop_state=None,
                 maxiters=10_000):
        self.args = (x0, qL, lin, means, chols)
        N, dev = means.shape[0], means.device
        if chunk_len is None:
            chunk_len = nat.default_chunk_len(N, lin["d"], lin["q"], dev.index)
        # the graph bakes the workspace address in: this object owns the workspace for as long as it lives
        self.ws = nat.Workspace(N, lin["d"], lin["q"], chunk_len, dev, means.dtype)
        self.kw = dict(calibrate=calibrate, chunk_len=chunk_len, scalars=scalars, ws=self.ws, loop_state=loop_state,
                       maxiters=maxiters)
        self.graph = None
        self._loop = None

    def _eager(self):
        run_iteration(*self.args, **self.kw)

    def capture_loop(self):
        """Build the WHOLE loop as one CUDA graph (`pof_ieks_loop_create`: the body in a WHILE conditional node whose
        condition the stopping-rule kernel sets on the device).  Needs `loop_state`; call after one eager iteration.
        Returns False if the driver cannot build it (the caller then replays single iterations)."""
        x0, qL, lin, means, chols = self.args
        kw = self.kw
        if kw["loop_state"] is None:
            raise nat.NativeError("capture_loop needs a loop_state")
        ivp_id, params = lin["builtin"]
        ph, pp = nat.host_doubles(list(params) + [0.0])
        qLh, qLp = nat.host_doubles(qL)
        out = ctypes.c_void_p()
        rc = nat.fn("pof_ieks_loop_create", means.dtype)(
            ctypes.byref(out), self.ws.ctx.ptr, nat.flags(), ivp_id, pp, len(params), means.shape[0], lin["d"],
            lin["q"], int(kw["chunk_len"]), qLp, lin["scale0"], lin["scale1"], nat.ptr(x0.mean), nat.ptr(x0.chol),
            nat.ptr(means), nat.ptr(chols), int(bool(kw["calibrate"])), nat.ptr(kw["scalars"]),
            nat.ptr(kw["loop_state"]), int(kw["maxiters"]), self.ws.ws_ptr, self.ws.nbytes)
        if rc != 0:
            return False
        self._loop = out
        return True

    def launch_loop(self):
        """enqueue the loop graph on the current stream: iterates until the device-side rule stops it"""
        nat.check(nat.LIB.pof_ieks_loop_launch(self._loop, nat.stream_ptr()), "pof_ieks_loop_launch")

    def __del__(self):
        try:
            if self._loop:
                nat.LIB.pof_ieks_loop_destroy(self._loop)
                self._loop = None
        except Exception:
            pass

    def capture(self):
        """call after at least one eager iteration (kernel attributes and workspaces exist)"""
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = g

    def __call__(self):
        if self.graph is None:
            self._eager()
        else:
            self.graph.replay()


class GraphedCall:
    """Any stream-ordered, host-sync-free callable (e.g. one time-sharded iteration: linearise + ShardedPass.run with
    its NCCL all-gathers) captured once into a CUDA graph and replayed.  `fn` must work on fixed buffers; its return
    value (device tensors) is kept from the capture and returned by every replay."""

    def __init__(self, fn):
        self.fn = fn
        self.graph = None
        self.out = None

    def capture(self):
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            with torch.cuda.graph(g, stream=side):
                self.out = self.fn()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = g

    def __call__(self):
        if self.graph is None:
            return self.fn()
        self.graph.replay()
        return self.out


def _outside_the_kernels(dom: AffineModel):
    """True for model shapes the CUDA kernels do not take: observation dimension != ODE dimension d with D = d (q+1),
    or q = 0 (e.g. the 1-dimensional Wiener process of the reference's tests/test_filtsmooth.py).  Those run through
    torch's batched library calls on the tensors' device (`library_pass`), announced by a warning -- never silently."""
    _, dy, D = dom.H.shape
    out = D % dy != 0 or D // dy < 2
    if out:
        import warnings

        warnings.warn(f"pof: observation dimension {dy} with state dimension {D} is outside the CUDA kernels "
                      "(they need D = d (q+1), q >= 1): this pass runs through torch library calls", stacklevel=3)
    return out


def smoothing(transition_models, filtered_states):
    """reference parallel_filtsmooth/smoother.py:8-22 as a stand-alone call: RTS smoothing of GIVEN filtered states by
    the associative suffix scan -> (MVNSqrt smoothed, obj).  The kernels of this package never materialise this seam
    (the filter scan hands its backward kernels straight to the smoother inside `linear_filtsmooth`), so the
    stand-alone form runs through torch's batched library calls on the states' device (`library_pass`)."""
    from .library_pass import smoothing_library

    return smoothing_library(transition_models, filtered_states)


def linear_filtsmooth(x0, linear_transitions, linear_observations, *, chunk_len=None, info=None):
    """reference parallel_filtsmooth/__init__.py:5-10 -> (MVNSqrt(means (N,D), chols (N,D,D)), nll, obj, ssq).
    `info` (optional dict) receives the pass's full scalar vector (`info["scalars"]`, indices `_native.S_*`, e.g. the
    QR-sign-invariant sigma^2 `S_SSQ_PROPER` that the reference does not return)."""
    if _outside_the_kernels(linear_observations):
        from .library_pass import linear_filtsmooth_library

        return linear_filtsmooth_library(x0, linear_transitions, linear_observations)
    n, d, q, D, qL, dense = _model_dims(linear_transitions, linear_observations)
    dev = linear_observations.H.device
    means = torch.zeros((n + 1, D), dtype=torch.float64, device=dev)
    chols = torch.empty((n + 1, D, D), dtype=torch.float64, device=dev)
    sc = run_pass(x0, qL, linear_observations.H.contiguous(), linear_observations.b.contiguous(), means, chols, d=d,
                  q=q, calibrate=False, chunk_len=chunk_len, cholR=_noise(linear_observations), dense=dense)
    if info is not None:
        info["scalars"] = sc
    return MVNSqrt(means, chols), sc[nat.S_NLL], sc[nat.S_OBJ], sc[nat.S_SSQ]


def linear_noiseless_filtering(x0, transition_models, observation_models, *, chunk_len=None):
    """reference parallel_filtsmooth/filter.py:18-47 -> (filtered MVNSqrt, nll, obj, ssq).
    The filtered chols are square-root factors (chol @ chol.T is the covariance) but not triangular."""
    if _outside_the_kernels(observation_models):
        from .library_pass import linear_noiseless_filtering_library

        return linear_noiseless_filtering_library(x0, transition_models, observation_models)
    n, d, q, D, qL, dense = _model_dims(transition_models, observation_models)
    dev = observation_models.H.device
    means = torch.zeros((n + 1, D), dtype=torch.float64, device=dev)
    fm = torch.empty((n + 1, D), dtype=torch.float64, device=dev)
    fc = torch.empty((n + 1, D, D), dtype=torch.float64, device=dev)
    sc = run_pass(x0, qL, observation_models.H.contiguous(), observation_models.b.contiguous(), means, None, d=d, q=q,
                  calibrate=False, chunk_len=chunk_len, fmeans=fm, fchols=fc, cholR=_noise(observation_models),
                  dense=dense)
    # filter objective (reference filter.py:43-45, swapped-argument form), evaluated on the filtered means
    F, QL = transition_models.F, transition_models.QL
    Fm = fm[1:] @ F.T if F.dim() == 2 else torch.einsum("nij,nj->ni", F, fm[1:])
    r = torch.linalg.solve_triangular(QL, (fm[:-1] - Fm).unsqueeze(-1), upper=False)
    obj = (r * r).sum()
    return MVNSqrt(fm, fc), sc[nat.S_NLL], obj, sc[nat.S_SSQ]


def _pack(parts):
    n = parts[0].shape[0]
    return torch.cat([p.reshape(n, -1) for p in parts], dim=1).contiguous()


def sqrt_filtering_operator(elem1, elem2):
    """reference filter.py:117-142 (batched over the leading axis): elems are tuples (A, b, U, eta, Z)"""
    n, D = elem1[1].shape
    e1, e2 = _pack(elem1), _pack(elem2)
    nat.require_cuda(e1, e2)
    out = torch.empty_like(e1)
    nat.check(nat.LIB.pof_filter_combine_f64(nat.stream_ptr(), n, D, nat.ptr(e1), nat.ptr(e2), nat.ptr(out),
                                             nat.flags()),
              "pof_filter_combine_f64")
    DD = D * D
    return (out[:, :DD].reshape(n, D, D), out[:, DD:DD + D], out[:, DD + D:2 * DD + D].reshape(n, D, D),
            out[:, 2 * DD + D:2 * DD + 2 * D], out[:, 2 * DD + 2 * D:].reshape(n, D, D))


def sqrt_smoothing_operator(elem1, elem2):
    """reference smoother.py:53-63 (batched): elems are tuples (g, E, D); elem1 is the LATER element"""
    n, D = elem1[0].shape
    e1, e2 = _pack(elem1), _pack(elem2)
    nat.require_cuda(e1, e2)
    out = torch.empty_like(e1)
    nat.check(nat.LIB.pof_smooth_combine_f64(nat.stream_ptr(), n, D, nat.ptr(e1), nat.ptr(e2), nat.ptr(out),
                                             nat.flags()),
              "pof_smooth_combine_f64")
    DD = D * D
    return out[:, :D], out[:, D:D + DD].reshape(n, D, D), out[:, D + DD:].reshape(n, D, D)
