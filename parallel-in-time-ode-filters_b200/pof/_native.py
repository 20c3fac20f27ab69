"""ctypes binding of libpof_b200.so (the C ABI declared in include/pof_b200.h).

The product path has NO CPU fallback: importing this module without the built CUDA library raises.
Device memory, streams and (for time-sharded runs) collectives are torch's; every numerical step of the hot path is
a kernel of the library.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POF_B200_LIB") or os.path.join(os.path.dirname(_HERE), "libpof_b200.so")

S_NLL, S_OBJ, S_SSQ, S_SSQ_PROPER, S_NOT_CLOSE, S_CSCALE, NSCALARS = 0, 1, 2, 3, 4, 5, 8

IVP_IDS = dict(
    logistic=0, lotkavolterra=1, vanderpol=2, fitzhughnagumo=3, rober=4, rigid_body=5, seir=6, threebody=7,
    henonheiles=8, lorenz96=9,
)

_c_dp = ctypes.c_void_p
_c_i64 = ctypes.c_int64
_c_int = ctypes.c_int
_c_dbl = ctypes.c_double
_c_sz = ctypes.c_size_t


class NativeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()').  There is no CPU fallback."
        )
    lib = ctypes.CDLL(LIB_PATH)
    U = ctypes.c_uint32
    sig = {
        "pof_supported": (_c_int, [_c_int, _c_int]),
        "pof_supported_tile": (_c_int, [_c_int, _c_int]),
        "pof_ctx_create": (_c_int, [ctypes.POINTER(ctypes.c_void_p)]),
        "pof_ctx_destroy": (None, [_c_dp]),
        "pof_ctx_profile_enable": (None, [_c_dp, _c_int]),
        "pof_ctx_profile_read": (_c_int, [_c_dp, _c_dp, _c_dp]),
        "pof_launches_per_pass": (_c_i64, [_c_i64, _c_int, _c_int, _c_i64, U]),
        "pof_measure_dfma_tflops": (_c_int, [_c_dp, _c_dp]),
        "pof_default_chunk_len": (_c_i64, [_c_i64, _c_int, _c_int, _c_int, U]),
        "pof_workspace_bytes": (_c_sz, [_c_i64, _c_int, _c_int, _c_i64]),
        "pof_filter_combine_f64": (_c_int, [_c_dp, _c_i64, _c_int, _c_dp, _c_dp, _c_dp, U]),
        "pof_smooth_combine_f64": (_c_int, [_c_dp, _c_i64, _c_int, _c_dp, _c_dp, _c_dp, U]),
        "pof_linearize_ivp_f64": (
            _c_int, [_c_dp, _c_int, _c_dp, _c_int, _c_i64, _c_int, _c_int, _c_dbl, _c_dbl, _c_dp, _c_dp, _c_dp]),
        "pof_linearize_ivp_compact_f64": (
            _c_int, [_c_dp, _c_int, _c_dp, _c_int, _c_i64, _c_int, _c_int, _c_dbl, _c_dp, _c_dp]),
        # (stream, ctx, flags, N, d, q, chunk_len, qL, x0m, x0c, H, c, means, chols, fmeans, fchols, calibrate,
        #  scalars, ws, ws_bytes)
        "pof_linear_filtsmooth_f64": (
            _c_int, [_c_dp, _c_dp, U, _c_i64, _c_int, _c_int, _c_i64, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp,
                     _c_dp, _c_dp, _c_int, _c_dp, _c_dp, _c_sz]),
        "pof_linear_filtsmooth_general_f64": (
            _c_int, [_c_dp, _c_dp, U, _c_i64, _c_int, _c_int, _c_i64, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp,
                     _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_int, _c_dp, _c_dp, _c_sz]),
        "pof_ieks_iteration_f64": (
            _c_int, [_c_dp, _c_dp, U, _c_int, _c_dp, _c_int, _c_i64, _c_int, _c_int, _c_i64, _c_dp, _c_dbl, _c_dbl,
                     _c_dp, _c_dp, _c_dp, _c_dp, _c_int, _c_dp, _c_dp, _c_sz]),
        "pof_ieks_loop_step_f64": (
            _c_int, [_c_dp, _c_dp, U, _c_int, _c_dp, _c_int, _c_i64, _c_int, _c_int, _c_i64, _c_dp, _c_dbl, _c_dbl,
                     _c_dp, _c_dp, _c_dp, _c_dp, _c_int, _c_dp, _c_dp, _c_i64, _c_dp, _c_sz]),
        # (out, ctx, flags, ivp, params, nparams, N, d, q, chunk_len, qL, s0, s1, x0m, x0c, means, chols, calibrate,
        #  scalars, loop_state, maxiters, ws, ws_bytes)
        "pof_ieks_loop_create_f64": (
            _c_int, [ctypes.POINTER(ctypes.c_void_p), _c_dp, U, _c_int, _c_dp, _c_int, _c_i64, _c_int, _c_int, _c_i64,
                     _c_dp, _c_dbl, _c_dbl, _c_dp, _c_dp, _c_dp, _c_dp, _c_int, _c_dp, _c_dp, _c_i64, _c_dp, _c_sz]),
        "pof_ieks_loop_launch": (_c_int, [_c_dp, _c_dp]),
        "pof_ieks_loop_destroy": (None, [_c_dp]),
        "pof_sequential_eks_f64": (
            _c_int, [_c_dp, U, _c_int, _c_dp, _c_int, _c_i64, _c_int, _c_int, _c_dp, _c_dbl, _c_dbl, _c_dp, _c_dp,
                     _c_dp, _c_dp, _c_dp, _c_dp, _c_sz]),
        "pof_shard_stage_a_f64": (
            _c_int, [_c_dp, _c_dp, U, _c_i64, _c_int, _c_int, _c_i64, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_sz]),
        "pof_shard_stage_b_f64": (
            _c_int, [_c_dp, _c_dp, U, _c_i64, _c_int, _c_int, _c_i64, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp,
                     _c_dp, _c_dp, _c_dp, _c_sz]),
        "pof_shard_stage_a_compact_f64": (
            _c_int, [_c_dp, _c_dp, U, _c_i64, _c_int, _c_int, _c_i64, _c_dp, _c_dp, _c_dbl, _c_dbl, _c_dp, _c_dp,
                     _c_sz]),
        "pof_shard_stage_b_compact_f64": (
            _c_int, [_c_dp, _c_dp, U, _c_i64, _c_int, _c_int, _c_i64, _c_dp, _c_dp, _c_dbl, _c_dbl, _c_dp, _c_dp,
                     _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_sz]),
        "pof_shard_stage_c_f64": (
            _c_int, [_c_dp, _c_dp, U, _c_i64, _c_int, _c_int, _c_i64, _c_dp, _c_dp, _c_int, _c_int, _c_dp, _c_dp,
                     _c_dp, _c_dp, _c_dp, _c_sz]),
        "pof_shard_exchange_supported": (_c_int, [_c_int, U]),
        "pof_shard_exchange_filter_f64": (
            _c_int, [_c_dp, U, _c_int, _c_int, _c_int, _c_dp, _c_i64, _c_dp, _c_dp, _c_dp, _c_dp]),
        "pof_shard_exchange_smooth_f64": (
            _c_int, [_c_dp, U, _c_int, _c_int, _c_int, _c_int, _c_i64, _c_int, _c_dp, _c_i64, _c_dp, _c_dp, _c_dp,
                     _c_dp]),
        "pof_shard_exchange_scalars_f64": (_c_int, [_c_dp, _c_int, _c_dp, _c_dp]),
        "pof_p2p_create": (_c_int, [_c_int, _c_int, _c_int, ctypes.POINTER(ctypes.c_void_p), _c_dp]),
        "pof_p2p_connect": (_c_int, [_c_dp, _c_dp]),
        "pof_p2p_destroy": (None, [_c_dp]),
        "pof_p2p_status": (_c_int, [_c_dp, ctypes.POINTER(ctypes.c_int)]),
        "pof_p2p_exchange_filter_f64": (_c_int, [_c_dp, U, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp]),
        "pof_p2p_exchange_smooth_f64": (
            _c_int, [_c_dp, U, _c_dp, _c_int, _c_i64, _c_int, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp]),
        "pof_p2p_exchange_scalars_f64": (_c_int, [_c_dp, _c_dp, _c_dp, _c_dp]),
        "pof_filter_apply_chain_f64": (_c_int, [_c_dp, U, _c_int, _c_int, _c_dp, _c_dp, _c_dp, _c_dp]),
        "pof_smooth_apply_chain_f64": (_c_int, [_c_dp, U, _c_int, _c_int, _c_dp, _c_dp, _c_dp, _c_dp]),
        "pof_project_f64": (_c_int, [_c_dp, _c_i64, _c_int, _c_int, _c_dbl, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp]),
        "pof_prior_init_f64": (_c_int, [_c_dp, _c_i64, _c_int, _c_int, _c_dp, _c_dp, _c_dp, _c_dp, _c_dp]),
    }
    # the optional fp32 mode: the register-resident kernels compiled with the scalar type float export the same entry
    # points with the suffix _f32 (same argument lists; device arrays are float, host-side arguments stay double)
    for name in F32_ENTRY_POINTS:
        sig[name] = sig[name[:-4] + "_f64"] if name != "pof_workspace_bytes_f32" else sig["pof_workspace_bytes"]
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return lib


F32_ENTRY_POINTS = [
    "pof_workspace_bytes_f32", "pof_filter_combine_f32", "pof_smooth_combine_f32", "pof_linearize_ivp_f32",
    "pof_linearize_ivp_compact_f32", "pof_linear_filtsmooth_f32", "pof_ieks_iteration_f32", "pof_ieks_loop_step_f32",
    "pof_ieks_loop_create_f32",
    "pof_shard_stage_a_f32",
    "pof_shard_stage_b_f32", "pof_shard_stage_a_compact_f32", "pof_shard_stage_b_compact_f32", "pof_shard_stage_c_f32",
    "pof_shard_exchange_filter_f32", "pof_shard_exchange_smooth_f32", "pof_shard_exchange_scalars_f32",
    "pof_prior_init_f32", "pof_project_f32",
]
LIB = _load()
EXPORTED = [
    "pof_supported", "pof_supported_tile", "pof_ctx_create", "pof_ctx_destroy", "pof_ctx_profile_enable",
    "pof_ctx_profile_read", "pof_launches_per_pass", "pof_measure_dfma_tflops", "pof_default_chunk_len",
    "pof_workspace_bytes", "pof_filter_combine_f64", "pof_smooth_combine_f64", "pof_linearize_ivp_f64",
    "pof_linearize_ivp_compact_f64", "pof_linear_filtsmooth_f64", "pof_linear_filtsmooth_general_f64",
    "pof_ieks_iteration_f64", "pof_ieks_loop_step_f64", "pof_ieks_loop_create_f64", "pof_ieks_loop_launch",
    "pof_ieks_loop_destroy", "pof_sequential_eks_f64", "pof_shard_stage_a_f64", "pof_shard_stage_b_f64",
    "pof_shard_stage_a_compact_f64", "pof_shard_stage_b_compact_f64", "pof_shard_stage_c_f64",
    "pof_filter_apply_chain_f64", "pof_smooth_apply_chain_f64", "pof_project_f64", "pof_prior_init_f64",
    "pof_shard_exchange_supported", "pof_shard_exchange_filter_f64", "pof_shard_exchange_smooth_f64",
    "pof_shard_exchange_scalars_f64", "pof_p2p_create", "pof_p2p_connect", "pof_p2p_destroy", "pof_p2p_status",
    "pof_p2p_exchange_filter_f64", "pof_p2p_exchange_smooth_f64", "pof_p2p_exchange_scalars_f64",
] + F32_ENTRY_POINTS
P2P_HANDLE_BYTES = 64


def fn(base, dtype):
    """entry point `base`_f64 or `base`_f32 for a tensor dtype (fp32: the register-resident family only)"""
    if dtype == torch.float64:
        return getattr(LIB, base + "_f64")
    if dtype == torch.float32:
        return getattr(LIB, base + "_f32")
    raise NativeError(f"unsupported dtype {dtype}")

# kernel-family flags of the C ABI (include/pof_b200.h).  DEFAULT_FLAGS is what the facade passes; tests / scripts may
# change it (e.g. F_FAMILY_TILE to run the large-state kernels on small problems) -- the library itself holds no
# switches and reads no environment variables.
F_FAMILY_TILE, F_TILE_SMEM_QR, F_TREE_PER_LEVEL, F_SMOOTH_TMA, F_TREE_UPDOWN = 1, 2, 4, 8, 16
DEFAULT_FLAGS = int(os.environ.get("POF_B200_FLAGS", "0"))
# solve(): run the IEKS loop as one CUDA graph with a WHILE conditional node (False: replay single iterations in bursts)
USE_LOOP_GRAPH = os.environ.get("POF_B200_LOOP_GRAPH", "1") != "0"
SEGMENTS = ["fold", "filter_up_sharded", "filter_tree", "scan", "smooth_up_side_stream", "smooth_down", "smooth"]


def flags(extra=0):
    return ctypes.c_uint32(int(DEFAULT_FLAGS) | int(extra))


def check(rc, what):
    if rc != 0:
        if rc > 0:
            raise NativeError(f"{what}: CUDA error {rc}")
        names = {-1: "unsupported (d, q)", -2: "workspace too small", -3: "bad argument", -4: "unknown IVP"}
        raise NativeError(f"{what}: {names.get(rc, rc)}")


def require_cuda(*tensors):
    dt = None
    for t in tensors:
        if t is None:
            continue
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype in (torch.float64, torch.float32)
                and t.is_contiguous()):
            raise NativeError("pof_b200 kernels need contiguous float64 (or, fp32 mode, float32) CUDA tensors "
                              "(there is no CPU fallback)")
        if dt is not None and t.dtype != dt:
            raise NativeError("all arrays of a call must have the same dtype")
        dt = t.dtype


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def host_doubles(a):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return a, a.ctypes.data_as(ctypes.c_void_p)


_SM_COUNT = {}


def sm_count(device=None):
    dev = torch.cuda.current_device() if device is None else device
    if dev not in _SM_COUNT:
        _SM_COUNT[dev] = torch.cuda.get_device_properties(dev).multi_processor_count
    return _SM_COUNT[dev]


class Context:
    """Caller-owned execution context of the C ABI (pof_ctx_t): the side stream on which a pass runs the smoother's
    up-sweep concurrently with the filter scan, plus optional per-segment timing."""

    def __init__(self, device=None):
        self._p = ctypes.c_void_p()
        if device is not None and torch.cuda.is_available():
            with torch.cuda.device(device):
                check(LIB.pof_ctx_create(ctypes.byref(self._p)), "pof_ctx_create")
        else:
            check(LIB.pof_ctx_create(ctypes.byref(self._p)), "pof_ctx_create")

    @property
    def ptr(self):
        return self._p

    def profile_enable(self, on=True):
        LIB.pof_ctx_profile_enable(self._p, int(bool(on)))

    def profile_read(self):
        """-> dict segment name -> (accumulated ms, count); synchronises the device"""
        ms = (ctypes.c_double * len(SEGMENTS))()
        cnt = (ctypes.c_int64 * len(SEGMENTS))()
        check(LIB.pof_ctx_profile_read(self._p, ms, cnt), "pof_ctx_profile_read")
        return {nm: (ms[i], cnt[i]) for i, nm in enumerate(SEGMENTS)}

    def __del__(self):
        try:
            if self._p:
                LIB.pof_ctx_destroy(self._p)
                self._p = ctypes.c_void_p()
        except Exception:
            pass


class Workspace:
    """Caller-owned scratch memory for one problem shape (N, d, q, chunk_len) on one device.

    Ownership rules (a pass writes GBs into this buffer, so a dangling pointer is silent corruption):
      * anything that bakes `buf.data_ptr()` into a CUDA graph (GraphedIteration, GraphedCall users) or keeps using it
        across calls (CudaBackend) OWNS its Workspace object -- create it with `Workspace(...)` and keep the reference;
      * `Workspace.get` is a small per-(shape, device, stream) cache for one-off eager calls.  Two passes on different
        streams never share a buffer; eviction only drops the cache's own reference (torch's stream-ordered caching
        allocator keeps the memory valid for kernels already queued on the stream that used it).
    """

    _cache = {}
    _CACHE_MAX = 4

    def __init__(self, N, d, q, chunk_len, device, dtype=torch.float64):
        self.N, self.d, self.q, self.chunk_len = int(N), int(d), int(q), int(chunk_len)
        self.dtype = dtype
        wsb = LIB.pof_workspace_bytes if dtype == torch.float64 else LIB.pof_workspace_bytes_f32
        self.nbytes = int(wsb(self.N, self.d, self.q, self.chunk_len))
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.ctx = Context(device)  # side stream + events of the passes that use this workspace

    @property
    def ws_ptr(self):
        return ctypes.c_void_p(self.buf.data_ptr())

    def matches(self, N, d, q, chunk_len, dtype=torch.float64):
        return (self.N, self.d, self.q, self.chunk_len, self.dtype) == (int(N), int(d), int(q), int(chunk_len), dtype)

    @classmethod
    def get(cls, N, d, q, chunk_len, device, dtype=torch.float64):
        stream = torch.cuda.current_stream(device).cuda_stream if torch.cuda.is_available() else 0
        key = (int(N), int(d), int(q), int(chunk_len), str(device), int(stream), str(dtype))
        ws = cls._cache.pop(key, None)
        if ws is None:
            while len(cls._cache) >= cls._CACHE_MAX:
                cls._cache.pop(next(iter(cls._cache)))  # least recently used; holders keep their own reference
            ws = cls(N, d, q, chunk_len, device, dtype)
        cls._cache[key] = ws  # most recently used last
        return ws


def default_chunk_len(N, d, q, device=None, extra_flags=0):
    return int(LIB.pof_default_chunk_len(int(N), int(d), int(q), sm_count(device), flags(extra_flags)))


def default_chunk_len_tile(N, d, q, device=None):
    """chunk length for the large-state (CTA-per-chunk) kernels, which also serve noisy observations"""
    return default_chunk_len(N, d, q, device, F_FAMILY_TILE)
