"""Time-sharded (multi-GPU) IEKS pass: one process per GPU, contiguous time shards, two carry exchanges.

Rank r owns steps [k_lo, k_hi) of the n = N-1 global steps and the state rows t in (k_lo, k_hi] (rank 0 also owns
row 0).  One pass is three local stages of the CUDA library with two exchange points:

    stage A  fold + up-sweep                 -> this shard's filtering element (3D^2+2D doubles)
      all-gather of the W filtering elements; every rank folds elements 0..r-1 onto x0 (<= W-1 combines, one warp)
    stage B  down-sweep, filter scan, smoother up-sweep
                                             -> smoothing element (2D^2+D), filtered end state (D+D^2), 3 partial sums
      all-gather of [smoothing element | end state | partial sums]; every rank folds the later ranks' elements onto the
      last rank's end state; the sums give nll / sigma^2 (hence the calibration scale) on every rank identically
    stage C  smoother down-sweep + smoother scan -> smoothed rows, 2 partial sums (objective, means not close)
      all-gather of the 2 sums

The exchanged payloads are a few KB, i.e. latency-bound: there is no bulk data-path collective.  The collectives are
`torch.distributed` (NCCL over NVLink on the GPUs; gloo in the CPU tests, where the stages run in the host simulator
through the same orchestration code).
"""
from __future__ import annotations

import ctypes
import math
import os

import torch
import torch.distributed as dist

from . import _native as nat


def shard_bounds(n_steps: int, rank: int, world: int):
    """contiguous step range [k_lo, k_hi) of `rank`"""
    return (rank * n_steps) // world, ((rank + 1) * n_steps) // world


class CudaBackend:
    """Stages of libpof_b200.so on the current CUDA device."""

    def __init__(self, d, q, n_loc, chunk_len, qL, device):
        self.d, self.q, self.n_loc = d, q, n_loc
        self.device = device
        self.chunk_len = int(chunk_len or nat.default_chunk_len(n_loc + 1, d, q, device.index))
        self.ws = nat.Workspace(n_loc + 1, d, q, self.chunk_len, device)  # owned: graph replays bake its address in
        self.qL, self.qLp = nat.host_doubles(qL)
        D = d * (q + 1)
        self.scratch = torch.empty(D + D * D, dtype=torch.float64, device=device)

    def _ws(self):
        return self.ws.ws_ptr, self.ws.nbytes

    def set_compact(self, scale0, scale1):
        """H arguments of stage_a/b are then the compact linearisation [J_f | c] (c must be None)"""
        self.scales = (float(scale0), float(scale1))

    def stage_a(self, H, c, carry_f):
        p, nb = self._ws()
        if c is None:
            s0, s1 = self.scales
            nat.check(nat.LIB.pof_shard_stage_a_compact_f64(nat.stream_ptr(), self.ws.ctx.ptr, nat.flags(), self.n_loc, self.d, self.q,
                                                            self.chunk_len, self.qLp, nat.ptr(H), s0, s1,
                                                            nat.ptr(carry_f), p, nb), "stage_a")
            return
        nat.check(nat.LIB.pof_shard_stage_a_f64(nat.stream_ptr(), self.ws.ctx.ptr, nat.flags(), self.n_loc, self.d, self.q, self.chunk_len, self.qLp,
                                                nat.ptr(H), nat.ptr(c), nat.ptr(carry_f), p, nb), "stage_a")

    def stage_b(self, H, c, state_in, fmeans, fchols, carry_s, state_end, partials):
        p, nb = self._ws()
        if c is None:
            s0, s1 = self.scales
            nat.check(nat.LIB.pof_shard_stage_b_compact_f64(nat.stream_ptr(), self.ws.ctx.ptr, nat.flags(), self.n_loc, self.d, self.q,
                                                            self.chunk_len, self.qLp, nat.ptr(H), s0, s1,
                                                            nat.ptr(state_in), nat.ptr(fmeans), nat.ptr(fchols),
                                                            nat.ptr(carry_s), nat.ptr(state_end), nat.ptr(partials), p,
                                                            nb), "stage_b")
            return
        nat.check(nat.LIB.pof_shard_stage_b_f64(nat.stream_ptr(), self.ws.ctx.ptr, nat.flags(), self.n_loc, self.d, self.q, self.chunk_len, self.qLp,
                                                nat.ptr(H), nat.ptr(c), nat.ptr(state_in), nat.ptr(fmeans),
                                                nat.ptr(fchols), nat.ptr(carry_s), nat.ptr(state_end),
                                                nat.ptr(partials), p, nb), "stage_b")

    def stage_c(self, seed, is_last, has_row0, cscale, means, chols, partials2):
        p, nb = self._ws()
        nat.check(nat.LIB.pof_shard_stage_c_f64(nat.stream_ptr(), self.ws.ctx.ptr, nat.flags(), self.n_loc, self.d, self.q, self.chunk_len, self.qLp,
                                                nat.ptr(seed), int(is_last), int(has_row0), nat.ptr(cscale),
                                                nat.ptr(means), nat.ptr(chols), nat.ptr(partials2), p, nb), "stage_c")

    # ---- fused exchanges (one launch each: carry fold + scalar bookkeeping), register-resident family only
    @property
    def fused_exchange(self):
        D = self.d * (self.q + 1)
        return bool(nat.LIB.pof_shard_exchange_supported(D, nat.flags()))

    def exchange_filter(self, D, rank, world, gathered, stride, x0_mean, x0_chol, state_in):
        nat.check(nat.LIB.pof_shard_exchange_filter_f64(nat.stream_ptr(), nat.flags(), D, rank, world,
                                                        nat.ptr(gathered), stride, nat.ptr(x0_mean), nat.ptr(x0_chol),
                                                        nat.ptr(state_in), nat.ptr(self.scratch)), "exchange_filter")

    def exchange_smooth(self, D, d, rank, world, n_total, calibrate, gathered, stride, seed, cscale, scalars):
        nat.check(nat.LIB.pof_shard_exchange_smooth_f64(nat.stream_ptr(), nat.flags(), D, d, rank, world, n_total,
                                                        int(bool(calibrate)), nat.ptr(gathered), stride, nat.ptr(seed),
                                                        nat.ptr(self.scratch), nat.ptr(cscale), nat.ptr(scalars)),
                  "exchange_smooth")

    def exchange_scalars(self, world, gathered, scalars):
        nat.check(nat.LIB.pof_shard_exchange_scalars_f64(nat.stream_ptr(), world, nat.ptr(gathered),
                                                         nat.ptr(scalars)), "exchange_scalars")

    def filter_chain(self, D, count, state_in, elems, state_out):
        nat.check(nat.LIB.pof_filter_apply_chain_f64(nat.stream_ptr(), nat.flags(), D, count, nat.ptr(state_in), nat.ptr(elems),
                                                     nat.ptr(state_out), nat.ptr(self.scratch)), "filter_chain")

    def smooth_chain(self, D, count, state_in, elems, state_out):
        nat.check(nat.LIB.pof_smooth_apply_chain_f64(nat.stream_ptr(), nat.flags(), D, count, nat.ptr(state_in), nat.ptr(elems),
                                                     nat.ptr(state_out), nat.ptr(self.scratch)), "smooth_chain")


class PeerExchange:
    """Peer-memory exchange areas of a time-sharded run (`pof_p2p_*`): one process per GPU of ONE NVLink / NVSwitch
    box.  Every rank allocates its area inside the library (CUDA IPC needs whole allocations), the 64-byte IPC handles
    are all-gathered once with torch.distributed, and from then on an exchange is ONE kernel per rank that stores its
    payload into every peer's area, releases a flag there, waits for the flags it needs and folds the carries -- no
    collective call in the pass."""

    def __init__(self, rank, world, D, group, device):
        import numpy as np

        self._p = ctypes.c_void_p()
        h = (ctypes.c_ubyte * nat.P2P_HANDLE_BYTES)()
        nat.check(nat.LIB.pof_p2p_create(rank, world, D, ctypes.byref(self._p), ctypes.cast(h, ctypes.c_void_p)),
                  "pof_p2p_create")
        mine = torch.tensor(list(h), dtype=torch.uint8, device=device)
        allh = torch.empty(world * nat.P2P_HANDLE_BYTES, dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(allh, mine, group=group)
        self._handles = np.ascontiguousarray(allh.cpu().numpy())
        nat.check(nat.LIB.pof_p2p_connect(self._p, self._handles.ctypes.data_as(ctypes.c_void_p)), "pof_p2p_connect")
        dist.barrier(group=group)  # every area is mapped everywhere before the first store

    @property
    def ptr(self):
        return self._p

    def status(self):
        st = ctypes.c_int(0)
        nat.check(nat.LIB.pof_p2p_status(self._p, ctypes.byref(st)), "pof_p2p_status")
        return st.value

    def close(self):
        if self._p:
            nat.LIB.pof_p2p_destroy(self._p)
            self._p = ctypes.c_void_p()


class ShardedPass:
    """One linear filter+smoother pass over a time-sharded trajectory (the multi-GPU form of
    pof.parallel_filtsmooth.linear_filtsmooth; reference pof/parallel_filtsmooth/__init__.py:5-10)."""

    def __init__(self, N, d, q, qL, *, rank=None, world=None, group=None, device=None, chunk_len=None, backend=None,
                 exchange="auto"):
        """exchange: "nccl" = all-gathers + one fused fold kernel per exchange; "p2p" = peer-memory exchange kernels
        (no collective in the pass; one NVLink box, register-resident kernel family); "auto" = p2p where it can be set
        up (world > 1, a real process group, CUDA IPC available), else nccl."""
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.N, self.d, self.q = int(N), int(d), int(q)
        self.D = D = d * (q + 1)
        self.n = self.N - 1
        self.k_lo, self.k_hi = shard_bounds(self.n, self.rank, self.world)
        self.n_loc = self.k_hi - self.k_lo
        if self.n_loc < 1:
            raise ValueError("every rank needs at least one time step")
        self.has_row0 = self.rank == 0
        self.rows = self.n_loc + (1 if self.has_row0 else 0)
        self.FE, self.SE, self.ST = 3 * D * D + 2 * D, 2 * D * D + D, D * D + D
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.backend = backend or CudaBackend(d, q, self.n_loc, chunk_len, qL, self.device)
        z = lambda *s: torch.zeros(*s, dtype=torch.float64, device=self.device)
        self.carry_f, self.gather_f = z(self.FE), z(self.world * self.FE)
        self.state_in, self.seed = z(self.ST), z(self.ST)
        # payload of the second exchange: [smoothing carry SE | end state ST | 3 partial sums | 1 pad]: an EVEN number of
        # doubles, so that every rank's slot in the gathered buffer stays 16-byte aligned (the kernels use 128-bit loads)
        self.PB = self.SE + self.ST + 4
        self.pay_b, self.gather_b = z(self.PB), z(self.world * self.PB)
        self.pay_c, self.gather_c = z(2), z(self.world * 2)
        self.cscale = z(1)
        self.x0_state = z(self.ST)
        self.scalars = z(nat.NSCALARS)
        self.p2p = None
        want = exchange if exchange != "auto" else os.environ.get("POF_B200_EXCHANGE", "auto")
        can = (self.world > 1 and backend is None and dist.is_available() and dist.is_initialized()
               and self.device.type == "cuda" and getattr(self.backend, "fused_exchange", False))
        if want in ("p2p", "auto") and can:
            try:
                self.p2p = PeerExchange(self.rank, self.world, D, group, self.device)
            except Exception:
                if want == "p2p":
                    raise
                self.p2p = None
            # every rank must take the same path: fall back together if any rank could not map its peers
            ok = torch.tensor([1 if self.p2p is not None else 0], dtype=torch.int32, device=self.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                if self.p2p is not None:
                    self.p2p.close()
                self.p2p = None
        self.exchange = "p2p" if self.p2p is not None else "nccl"

    def _all_gather(self, out, inp):
        if self.world == 1:
            out.copy_(inp)
        else:
            dist.all_gather_into_tensor(out, inp, group=self.group)

    def run(self, x0_mean, x0_chol, H_loc, c_loc, means_loc, chols_loc, *, calibrate=True, fmeans=None, fchols=None):
        """H_loc (n_loc,d,D), c_loc (n_loc,d): local linearisation; means_loc (rows,D) in/out; chols_loc (rows,D,D)
        out or None.  Returns dict(nll, obj, ssq, ssq_proper, not_close) -- identical on every rank."""
        D, W, r, be = self.D, self.world, self.rank, self.backend
        FE, SE, ST = self.FE, self.SE, self.ST
        if getattr(be, "fused_exchange", False):
            return self._run_fused(x0_mean, x0_chol, H_loc, c_loc, means_loc, chols_loc, calibrate, fmeans, fchols)
        self.x0_state[:D].copy_(x0_mean)
        self.x0_state[D:].copy_(x0_chol.reshape(-1))
        # ---- stage A + exchange 1
        be.stage_a(H_loc, c_loc, self.carry_f)
        self._all_gather(self.gather_f, self.carry_f)
        be.filter_chain(D, r, self.x0_state, self.gather_f, self.state_in)
        # ---- stage B + exchange 2
        pb = self.pay_b
        if fmeans is not None and self.has_row0:
            fmeans[0].copy_(x0_mean)
            fchols[0].copy_(x0_chol)
        shift = 0 if self.has_row0 else 1  # local row of state t' is t' - shift
        fm = None if fmeans is None else fmeans
        be.stage_b(H_loc, c_loc, self.state_in, _shifted(fm, shift, D), _shifted(fchols, shift, D * D), pb[:SE],
                   pb[SE:SE + ST], pb[SE + ST:SE + ST + 3])
        self._all_gather(self.gather_b, pb)
        gb = self.gather_b.view(W, self.PB)
        sums = gb[:, SE + ST:SE + ST + 3].sum(dim=0)  # same order on every rank -> bitwise identical scalars
        nll = sums[0]
        ssq = sums[1] / self.n / self.d
        ssq_proper = sums[2] / self.n / self.d
        self.cscale.copy_(torch.sqrt(ssq).reshape(1) if calibrate else torch.ones(1, dtype=torch.float64,
                                                                                   device=self.device))
        terminal = gb[W - 1, SE:SE + ST].contiguous()
        later = gb[r + 1:, :SE].contiguous() if r + 1 < W else gb[:0, :SE].contiguous()
        be.smooth_chain(D, W - 1 - r, terminal, later, self.seed)
        # ---- stage C + exchange 3
        be.stage_c(self.seed, r == W - 1, self.has_row0, self.cscale, means_loc, chols_loc, self.pay_c)
        self._all_gather(self.gather_c, self.pay_c)
        s2 = self.gather_c.view(W, 2).sum(dim=0)
        return dict(nll=nll, obj=s2[0], ssq=ssq, ssq_proper=ssq_proper, not_close=s2[1])


def _run_fused(self, x0_mean, x0_chol, H_loc, c_loc, means_loc, chols_loc, calibrate, fmeans, fchols):
    """the same pass with ONE kernel per exchange (carry fold + scalar bookkeeping, `pof_shard_exchange_*`): per pass
    three stage calls, three all-gathers and three exchange kernels -- no other device work.  Written as phases so
    that a test can drive several "virtual ranks" in lockstep on one GPU (tests/test_gpu_sharded.py)."""
    st = (x0_mean, x0_chol, H_loc, c_loc, means_loc, chols_loc, calibrate, fmeans, fchols)
    self.phase_a(st)
    if self.p2p is None:
        self._all_gather(self.gather_f, self.carry_f)
    self.phase_b(st)
    if self.p2p is None:
        self._all_gather(self.gather_b, self.pay_b)
    self.phase_c(st)
    if self.p2p is None:
        self._all_gather(self.gather_c, self.pay_c)
    return self.phase_d()


def _phase_a(self, st):
    self.backend.stage_a(st[2], st[3], self.carry_f)


def _phase_b(self, st):
    x0_mean, x0_chol, H_loc, c_loc, _, _, _, fmeans, fchols = st
    D, SE, ST, pb = self.D, self.SE, self.ST, self.pay_b
    if self.p2p is not None:  # push my carry to every peer, wait for the earlier ranks', fold: one kernel, no collective
        nat.check(nat.LIB.pof_p2p_exchange_filter_f64(nat.stream_ptr(), nat.flags(), self.p2p.ptr,
                                                      nat.ptr(self.carry_f), nat.ptr(x0_mean), nat.ptr(x0_chol),
                                                      nat.ptr(self.state_in), nat.ptr(self.backend.scratch)),
                  "p2p_exchange_filter")
    else:
        self.backend.exchange_filter(D, self.rank, self.world, self.gather_f, self.FE, x0_mean, x0_chol, self.state_in)
    if fmeans is not None and self.has_row0:
        fmeans[0].copy_(x0_mean)
        fchols[0].copy_(x0_chol)
    shift = 0 if self.has_row0 else 1
    self.backend.stage_b(H_loc, c_loc, self.state_in, _shifted(fmeans, shift, D), _shifted(fchols, shift, D * D),
                         pb[:SE], pb[SE:SE + ST], pb[SE + ST:SE + ST + 3])


def _phase_c(self, st):
    means_loc, chols_loc, calibrate = st[4], st[5], st[6]
    D, SE, ST, W, r = self.D, self.SE, self.ST, self.world, self.rank
    if self.p2p is not None:
        nat.check(nat.LIB.pof_p2p_exchange_smooth_f64(nat.stream_ptr(), nat.flags(), self.p2p.ptr, self.d, self.n,
                                                      int(bool(calibrate)), nat.ptr(self.pay_b), nat.ptr(self.seed),
                                                      nat.ptr(self.backend.scratch), nat.ptr(self.cscale),
                                                      nat.ptr(self.scalars)), "p2p_exchange_smooth")
    else:
        self.backend.exchange_smooth(D, self.d, r, W, self.n, calibrate, self.gather_b, self.PB, self.seed,
                                     self.cscale, self.scalars)
    self.backend.stage_c(self.seed, r == W - 1, self.has_row0, self.cscale, means_loc, chols_loc, self.pay_c)


def _phase_d(self):
    sc = self.scalars
    if self.p2p is not None:
        nat.check(nat.LIB.pof_p2p_exchange_scalars_f64(nat.stream_ptr(), self.p2p.ptr, nat.ptr(self.pay_c),
                                                       nat.ptr(sc)), "p2p_exchange_scalars")
    else:
        self.backend.exchange_scalars(self.world, self.gather_c, sc)
    return dict(nll=sc[nat.S_NLL], obj=sc[nat.S_OBJ], ssq=sc[nat.S_SSQ], ssq_proper=sc[nat.S_SSQ_PROPER],
                not_close=sc[nat.S_NOT_CLOSE], scalars=sc)


ShardedPass.phase_a, ShardedPass.phase_b, ShardedPass.phase_c, ShardedPass.phase_d = _phase_a, _phase_b, _phase_c, _phase_d
ShardedPass._run_fused = _run_fused


def _shifted(t, shift, width):
    """view of `t` whose row index is offset by `shift` rows (the stages index local states by t', rows by t'-shift);
    row -1 is never touched when shift == 1"""
    if t is None or shift == 0:
        return t
    return _RowShift(t, shift * width)


class _RowShift:
    """pointer-only wrapper: data_ptr() moved back by `back` doubles (used only to pass a base address to the C ABI)"""

    def __init__(self, t, back):
        self.t, self.back = t, back
        self.is_cuda, self.dtype = t.is_cuda, t.dtype

    def data_ptr(self):
        return self.t.data_ptr() - 8 * self.back

    def is_contiguous(self):
        return self.t.is_contiguous()


def solve_sharded(*, f, y0, ts, order, init="constant", calibrate=True, maxiters=10_000, group=None, chunk_len=None,
                  graph=True):
    """The time-sharded form of `pof.solver.solve` (reference solver.py:11-73): one
    process per GPU, rank r owns the contiguous rows `rows = slice(r0, k_hi + 1)` of the N grid points.  Every rank
    runs the same IEKS loop; per iteration it linearises its shard (built-in `pof.ivp` problems: fused CUDA kernel,
    compact form; any other `f`: torch.func autodiff on the device, dense (H, c)), runs
    `ShardedPass.run` (three stages, three small all-gathers) and evaluates the reference's stopping rule on scalars
    that are bitwise identical on all ranks -- so all ranks leave the loop in the same iteration without any further
    exchange.  After the first (eager) iteration the whole iteration is replayed from one CUDA graph per rank.

    Returns (MVNSqrt(mean (rows, d), chol (rows, d, D)) of the local rows, info_dict, rows)."""
    from .convenience import get_initial_trajectory, set_up_solver
    from .convergence_criteria import crit_scalars
    from .parallel_filtsmooth import GraphedCall
    from .utils import MVNSqrt

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    setup = set_up_solver(f=f, y0=y0, ts=ts, order=order)
    lin = setup["om"].f._pof_lin
    builtin = lin["builtin"] is not None
    x0, dev = setup["x0"], setup["_device"]
    d, q = lin["d"], order
    D = d * (q + 1)
    N = len(setup["ts"])
    sp = ShardedPass(N, d, q, setup["_qL"], rank=rank, world=world, group=group, device=dev, chunk_len=chunk_len)
    if builtin:
        sp.backend.set_compact(lin["scale0"], lin["scale1"])
    r0 = 0 if sp.has_row0 else sp.k_lo + 1
    rows = slice(r0, sp.k_hi + 1)
    full = get_initial_trajectory(setup, method=init, means_only=True)
    means = full.mean[rows].contiguous().clone()
    del full
    chols = torch.empty((sp.rows, D, D), dtype=torch.float64, device=dev)
    t1row = 1 if sp.has_row0 else 0
    out5 = torch.zeros(5, dtype=torch.float64, device=dev)
    if builtin:
        Jc = torch.empty((sp.n_loc, d * d + d), dtype=torch.float64, device=dev)
        ivp_id, params = lin["builtin"]
        ph, pp = nat.host_doubles(list(params) + [0.0])
    else:
        from .step import linearize_into

        H = torch.empty((sp.n_loc, d, D), dtype=torch.float64, device=dev)
        c = torch.empty((sp.n_loc, d), dtype=torch.float64, device=dev)
        graph = False  # autodiff of a user function is not captured into a CUDA graph

    def iteration():
        if builtin:
            nat.check(nat.LIB.pof_linearize_ivp_compact_f64(nat.stream_ptr(), ivp_id, pp, len(params), sp.n_loc, d, q,
                                                            lin["scale0"], nat.ptr(means[t1row:]), nat.ptr(Jc)),
                      "linearize")
            res = sp.run(x0.mean, x0.chol, Jc, None, means, chols, calibrate=True)  # the loop always calibrates
        else:
            # linearize_into linearises at rows 1.. of what it is given: prepend one (unused) row on ranks > 0
            pts = means if sp.has_row0 else torch.cat([means[:1], means])
            linearize_into(lin, pts, H, c)
            res = sp.run(x0.mean, x0.chol, H, c, means, chols, calibrate=True)
        if "scalars" in res:
            return res["scalars"]  # (nll, obj, ssq, ssq_proper, not_close, ...): written by the exchange kernels
        out5.copy_(torch.stack([res["nll"], res["obj"], res["ssq"], res["ssq_proper"], res["not_close"]]))
        return out5

    step = GraphedCall(iteration)
    nll = obj = ssq = ssqp = 0.0
    nll_old = obj_old = 0.0
    n_bad = 1.0
    k = 0
    while True:
        if k >= 1:
            if crit_scalars(obj, obj_old, nll, nll_old, n_bad) or not (k <= maxiters):
                break
        nll_old, obj_old = nll, obj
        if graph and step.graph is None and k >= 1 and dev.type == "cuda":
            step.capture()
        sc = step().cpu()
        nll, obj, ssq, ssqp, n_bad = (float(v) for v in sc[:5])
        k += 1
    info = {"iterations": k, "nll": nll, "obj": obj, "sigma_squared": ssq, "calibrated": bool(calibrate),
            "sigma_squared_proper": ssqp}
    # final calibration (the second one, solver.py:66-69) fused with the E0 projection (solver.py:71)
    ymean = torch.empty((sp.rows, d), dtype=torch.float64, device=dev)
    ychol = torch.empty((sp.rows, d, D), dtype=torch.float64, device=dev)
    mult = sp.cscale if calibrate else None  # sqrt(sigma^2) of the last pass, identical on every rank
    nat.check(nat.LIB.pof_project_f64(nat.stream_ptr(), sp.rows, d, q, setup["_scale0"], nat.ptr(mult),
                                      nat.ptr(means), nat.ptr(chols), nat.ptr(ymean), nat.ptr(ychol)), "project")
    step.graph = None  # release the captured NCCL work before the caller tears the process group down
    return MVNSqrt(ymean, ychol), info, rows
