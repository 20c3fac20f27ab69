"""Dry run of tests/test_gpu_tile.py WITHOUT a GPU (developer aid, not a test and not a product path).

The C ABI entry points the facade calls are replaced by a fake library that forwards to the HOST SIMULATOR of the tile
kernels (tests/hostsim: the same device code compiled for the CPU); torch tensors live on the CPU, CUDA-graph capture
and synchronisation are no-ops.  What this validates is everything ABOVE the kernels: the facade's routing and argument
marshalling (pointer order, shapes, chunk lengths, workspace sizes), and the test code itself -- so that a red GPU run
points at a kernel, not at Python.  Written when the tile kernels could not be run on a B200 (GPU budget spent).

    python scripts/dryrun_gpu_tests_on_host.py [-k expr]
    POF_DRYRUN_FILE=test_gpu_parity.py python scripts/dryrun_gpu_tests_on_host.py -k "tile"
        (the parity tests of the tile family through the same fake library)
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200"), os.path.join(ROOT, "tests")]
import numpy as np  # noqa: E402
import torch  # noqa: E402

HS = ctypes.CDLL(os.path.join(ROOT, "tests", "hostsim", "libhostsim.so"))
P = ctypes.c_void_p


def _arr(ptr, shape):
    """numpy view of `ptr` (a c_void_p / int / None) with the given shape of doubles"""
    if ptr is None:
        return None
    addr = ptr.value if isinstance(ptr, ctypes.c_void_p) else int(ptr)
    if not addr:
        return None
    n = int(np.prod(shape))
    buf = (ctypes.c_double * n).from_address(addr)
    return np.frombuffer(buf, dtype=np.float64).reshape(shape)


def _p(a):
    return None if a is None else a.ctypes.data_as(P)


def _reg(flags):
    """register-resident Householder sweeps unless POF_F_TILE_SMEM_QR (2) is set"""
    f = flags.value if hasattr(flags, "value") else int(flags)
    return 0 if (f & 2) else 1


class FakeLib:
    """the subset of libpof_b200.so the tile tests reach, on the host simulator"""

    def __init__(self, real):
        self.real = real
        self.calls = {}

    def __getattr__(self, name):  # host-only queries go to the real library
        return getattr(self.real, name)

    def _count(self, name):
        self.calls[name] = self.calls.get(name, 0) + 1

    def _pass(self, N, d, q, L, qL, x0m, x0c, H, c, Jc, s0, s1, R, F, QL, means, chols, fm, fc, calibrate, scalars,
              flags=0):
        D = d * (q + 1)
        n = N - 1
        qLa = np.ascontiguousarray(_arr(qL, (q + 1, q + 1))) if qL else np.zeros((q + 1, q + 1))
        x0 = np.concatenate([_arr(x0m, (D,)), _arr(x0c, (D, D)).ravel()])
        sc = np.zeros(8)
        rc = HS.hs_tile_linear_filtsmooth(
            d, q, ctypes.c_long(N), ctypes.c_long(int(L)), _p(qLa), _p(x0), H, c, Jc, ctypes.c_double(s0),
            ctypes.c_double(s1), R, F, QL, means, chols, fm, fc, int(calibrate), _p(sc), 0, _reg(flags))
        out = _arr(scalars, (8,))
        nll, obj, ssq, ssqp, bad = sc[:5]
        out[:] = 0.0
        out[0], out[1], out[2], out[3], out[4] = nll, obj, ssq, ssqp, bad
        out[5] = np.sqrt(ssq) if calibrate else 1.0
        assert n >= 1
        return rc

    def pof_linear_filtsmooth_f64(self, s, ctx, flags, N, d, q, L, qL, x0m, x0c, H, c, means, chols, fm, fc, calibrate,
                                  scalars, ws, ws_bytes):
        self._count("pof_linear_filtsmooth_f64")
        assert ws_bytes >= self.real.pof_workspace_bytes(N, d, q, L)
        return self._pass(N, d, q, L, qL, x0m, x0c, H, c, None, 0.0, 0.0, None, None, None, means, chols, fm, fc,
                          calibrate, scalars, flags)

    def pof_linear_filtsmooth_general_f64(self, s, ctx, flags, N, d, q, L, qL, F, QL, x0m, x0c, H, c, R, means, chols,
                                          fm, fc, calibrate, scalars, ws, ws_bytes):
        self._count("pof_linear_filtsmooth_general_f64")
        assert ws_bytes >= self.real.pof_workspace_bytes(N, d, q, L)
        return self._pass(N, d, q, L, qL, x0m, x0c, H, c, None, 0.0, 0.0, R, F, QL, means, chols, fm, fc, calibrate,
                          scalars, flags)

    def pof_linearize_ivp_f64(self, s, ivp_id, params, nparams, n, d, q, s0, s1, means_t1, H, c):
        self._count("pof_linearize_ivp_f64")
        D = d * (q + 1)
        m = _arr(means_t1, (n, D))
        Ha, ca = _arr(H, (n, d, D)), _arr(c, (n, d))
        if ivp_id == 9:
            forcing = _arr(params, (1,))[0]
            Jc = np.zeros((n, d * d + d))
            return HS.hs_linearize_l96(ctypes.c_double(forcing), ctypes.c_long(n), d, q, ctypes.c_double(s0),
                                       ctypes.c_double(s1), _p(np.ascontiguousarray(m)), H, c, _p(Jc))
        # the other built-ins: oracle vector fields (this harness is about the plumbing, not about k_linearize)
        from oracle import ivps as oivps

        names = ["logistic", "lotkavolterra", "vanderpol", "fitzhughnagumo", "rober", "rigid_body", "seir", "threebody",
                 "henonheiles"]
        pv = list(_arr(params, (max(nparams, 1),))[:nparams])
        kw = {}
        if names[ivp_id] == "vanderpol":
            kw = {"stiffness_constant": pv[0]}
        elif names[ivp_id] == "henonheiles":
            kw = {"p": pv[0]}
        elif names[ivp_id] not in ("logistic",):
            kw = {"p": tuple(pv)}
        oivp = getattr(oivps, names[ivp_id])(**kw)
        Q1 = q + 1
        for k in range(n):
            y = s0 * m[k, 0::Q1]
            J, f = oivp.jac(y), oivp.f(None, y)
            Ha[k] = 0.0
            for a in range(d):
                Ha[k, a, 0::Q1] = -J[a] * s0
                Ha[k, a, a * Q1 + 1] += s1
            ca[k] = J @ y - f
        return 0

    def pof_ieks_iteration_f64(self, s, ctx, flags, ivp_id, params, nparams, N, d, q, L, qL, s0, s1, x0m, x0c, means,
                               chols, calibrate, scalars, ws, ws_bytes):
        self._count("pof_ieks_iteration_f64")
        D = d * (q + 1)
        n = N - 1
        m = _arr(means, (N, D))
        if ivp_id != 9:  # other built-ins: dense linearisation through the fake entry point above, then the pass
            H, c = np.zeros((n, d, D)), np.zeros((n, d))
            m1 = np.ascontiguousarray(m[1:])
            self.pof_linearize_ivp_f64(s, ivp_id, params, nparams, n, d, q, s0, s1, _p(m1), _p(H), _p(c))
            return self._pass(N, d, q, L, qL, x0m, x0c, _p(H), _p(c), None, 0.0, 0.0, None, None, None, means, chols,
                              None, None, calibrate, scalars, flags)
        forcing = _arr(params, (1,))[0]
        H, c, Jc = np.zeros((n, d, D)), np.zeros((n, d)), np.zeros((n, d * d + d))
        HS.hs_linearize_l96(ctypes.c_double(forcing), ctypes.c_long(n), d, q, ctypes.c_double(s0), ctypes.c_double(s1),
                            _p(np.ascontiguousarray(m[1:])), _p(H), _p(c), _p(Jc))
        return self._pass(N, d, q, L, qL, x0m, x0c, None, None, _p(Jc), s0, s1, None, None, None, means, chols, None,
                          None, calibrate, scalars, flags)

    def pof_sequential_eks_f64(self, s, flags, ivp_id, params, nparams, N, d, q, qL, s0, s1, x0m, x0c, means, chols,
                               scalars, ws, ws_bytes):
        self._count("pof_sequential_eks_f64")
        D = d * (q + 1)
        p8 = np.zeros(8)
        p8[:nparams] = _arr(params, (max(nparams, 1),))[:nparams]
        x0 = np.concatenate([_arr(x0m, (D,)), _arr(x0c, (D, D)).ravel()])
        qLa = np.ascontiguousarray(_arr(qL, (q + 1, q + 1)))
        sums = np.zeros(8)
        rc = HS.hs_tile_seq_eks(d, q, ctypes.c_long(N), _p(qLa), ctypes.c_double(s0), ctypes.c_double(s1), ivp_id,
                                _p(p8), _p(x0), means, chols, _p(sums), 0, _reg(flags))
        out = _arr(scalars, (8,))
        out[:] = 0.0
        out[0], out[1], out[2], out[3] = -sums[0], sums[3], sums[1] / (N - 1) / d, sums[2] / (N - 1) / d
        out[5] = 1.0
        return rc

    def pof_shard_exchange_supported(self, D, flags):
        return 0  # no register-resident family in the host simulator: solve() keeps its host-side loop

    def pof_ctx_create(self, out):
        return 0

    def pof_ctx_destroy(self, c):
        return None

    def pof_prior_init_f64(self, s, N, d, q, qL, ts, m0, means, chols):
        raise NotImplementedError("prior init has no host simulation")

    def pof_project_f64(self, s, N, d, q, scale0, mult, means, chols, ymean, ychol):
        self._count("pof_project_f64")
        D, Q1 = d * (q + 1), q + 1
        m, L = _arr(means, (N, D)), _arr(chols, (N, D, D))
        mu = _arr(mult, (1,))[0] if mult is not None and (mult.value if isinstance(mult, P) else mult) else 1.0
        _arr(ymean, (N, d))[:] = scale0 * m[:, 0::Q1]
        yc = _arr(ychol, (N, d, D))
        if yc is not None:
            yc[:] = mu * scale0 * L[:, 0::Q1, :]
        return 0

    def pof_filter_combine_f64(self, s, n, D, e1, e2, out, flags=0):
        self._count("pof_filter_combine_f64")
        FE = 3 * D * D + 2 * D
        a, b, o = _arr(e1, (n, FE)), _arr(e2, (n, FE)), _arr(out, (n, FE))
        for i in range(n):
            HS.hs_tile_filter_combine(D, _p(np.ascontiguousarray(a[i])), _p(np.ascontiguousarray(b[i])),
                                      _p(o[i]), 0, 0)
        return 0

    def pof_smooth_combine_f64(self, s, n, D, e1, e2, out, flags=0):
        self._count("pof_smooth_combine_f64")
        SE = 2 * D * D + D
        a, b, o = _arr(e1, (n, SE)), _arr(e2, (n, SE)), _arr(out, (n, SE))
        for i in range(n):
            HS.hs_tile_smooth_combine(D, _p(np.ascontiguousarray(a[i])), _p(np.ascontiguousarray(b[i])),
                                      _p(o[i]), 0, 0)
        return 0


def main():
    import pytest

    import pof.convenience as conv
    import pof.parallel_filtsmooth as pfs
    from pof import _native as nat

    fake = FakeLib(nat.LIB)
    nat.LIB = fake
    nat.require_cuda = lambda *tensors: None
    nat.stream_ptr = lambda: ctypes.c_void_p(0)
    nat.sm_count = lambda device=None: 148
    conv._device = lambda: torch.device("cpu")
    torch.cuda.synchronize = lambda *a, **k: None
    pfs.GraphedIteration.capture = lambda self: None
    _as_tensor = torch.as_tensor

    def as_tensor_cpu(data, *a, **kw):
        if kw.get("device") == "cuda":
            kw["device"] = "cpu"
        return _as_tensor(data, *a, **kw)

    torch.as_tensor = as_tensor_cpu

    class Plugin:
        def pytest_collection_modifyitems(self, items):
            for it in items:  # run the gpu-marked tests here
                it.own_markers = [m for m in it.own_markers if m.name != "gpu"]

    target = os.environ.get("POF_DRYRUN_FILE", "test_gpu_tile.py")
    args = [os.path.join(ROOT, "tests", target), "-q", "-p", "no:cacheprovider"] + sys.argv[1:]
    rc = pytest.main(args, plugins=[Plugin()])
    print("native calls:", fake.calls)
    return rc


if __name__ == "__main__":
    sys.exit(main())
