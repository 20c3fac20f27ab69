"""Device code on the host: the exact per-thread / per-warp code of the CUDA library (thread-per-chunk leaves, warp
tree combines, chunking, tree schedule) compiled for the CPU and checked against the oracle.  This is the CPU-side
proof that the blocked algorithm (fold -> tree up/down-sweep -> seeded scan -> smoother tree -> seeded RTS) computes
what the reference's associative scans compute."""
import ctypes
import os

import numpy as np
import pytest

from oracle import ivps
from oracle import pof_oracle as O

HS = os.path.join(os.path.dirname(__file__), "hostsim", "libhostsim.so")
P = ctypes.c_void_p


@pytest.fixture(scope="module")
def lib(native_lib):
    return ctypes.CDLL(HS)


def _cov(L):
    return L @ np.swapaxes(L, -1, -2)


def _p(a):
    return None if a is None else a.ctypes.data_as(P)


@pytest.mark.parametrize("name,kw,N,q,L", [
    ("fitzhughnagumo", {}, 100, 3, 4), ("fitzhughnagumo", {}, 100, 3, 7), ("fitzhughnagumo", {}, 100, 3, 200),
    ("fitzhughnagumo", {}, 1024, 3, 16), ("logistic", {}, 64, 3, 5), ("logistic", {}, 64, 1, 5),
    ("rigid_body", {}, 256, 3, 8), ("henonheiles", {"tmax": 10.0}, 128, 2, 8), ("lotkavolterra", {}, 300, 2, 9),
])
def test_blocked_pass_matches_oracle(lib, name, kw, N, q, L):
    ivp = getattr(ivps, name)(**kw)
    ts = np.linspace(ivp.t0, ivp.tmax, N)
    setup = O.set_up_solver(ivp, ts, q)
    st = O.get_initial_trajectory(setup)
    dom = O.linearize_at(setup, st.mean[1:])
    d = setup["d"]
    D = d * (q + 1)
    qL = np.ascontiguousarray(O.preconditioned_discretize_1d(q)[1])
    x0 = np.concatenate([setup["x0"].mean, setup["x0"].chol.ravel()])
    means, chols = st.mean.copy(), np.zeros((N, D, D))
    fm, fc, sc = np.zeros((N, D)), np.zeros((N, D, D)), np.zeros(8)
    H, c = np.ascontiguousarray(dom.H), np.ascontiguousarray(dom.b)
    rc = lib.hs_linear_filtsmooth(d, q, ctypes.c_long(N), ctypes.c_long(L), _p(qL), _p(x0), _p(H), _p(c), _p(means),
                                  _p(chols), _p(fm), _p(fc), 0, _p(sc))
    assert rc == 0
    filt, nll, _, ssq, ssqp = O.linear_noiseless_filtering(setup["x0"], setup["dtm"], dom)
    out, obj = O.smoothing(setup["dtm"], filt)
    sm = np.abs(out.mean).max()
    assert np.abs(fm - filt.mean).max() <= 1e-9 * np.abs(filt.mean).max()
    assert np.abs(_cov(fc) - _cov(filt.chol)).max() <= 1e-10 * np.abs(_cov(filt.chol)).max()
    assert np.abs(means - out.mean).max() <= 1e-9 * sm
    assert np.abs(_cov(chols) - _cov(out.chol)).max() <= 1e-10 * np.abs(_cov(out.chol)).max()
    assert abs(sc[0] - nll) <= 1e-9 * abs(nll) and abs(sc[1] - obj) <= 1e-9 * abs(obj)
    assert abs(sc[3] - ssqp) <= 1e-9 * ssqp and abs(sc[2] - ssq) <= 1e-2 * ssq
    # convergence counter: number of mean entries that moved (old = initial trajectory)
    moved = np.sum(~(np.abs(st.mean - out.mean) <= 1e-8 + 1e-13 * np.abs(out.mean)))
    assert abs(sc[4] - moved) <= 2


def test_warp_combine_matches_oracle_operator(lib):
    rng = np.random.default_rng(3)
    D = 8
    mk = lambda: (rng.standard_normal((1, D, D)), rng.standard_normal((1, D)), np.tril(rng.standard_normal((1, D, D))),
                  rng.standard_normal((1, D)), np.tril(rng.standard_normal((1, D, D))))
    for _ in range(5):
        e1, e2 = mk(), mk()
        pack = lambda e: np.concatenate([x.reshape(-1) for x in e])
        out = np.zeros(3 * D * D + 2 * D)
        assert lib.hs_filter_combine(D, _p(pack(e1)), _p(pack(e2)), _p(out), 0) == 0
        ref = O.sqrt_filtering_operator(e1, e2)
        DD = D * D
        np.testing.assert_allclose(out[:DD].reshape(D, D), ref[0][0], atol=1e-10 * np.abs(ref[0]).max())
        np.testing.assert_allclose(out[DD:DD + D], ref[1][0], atol=1e-10 * np.abs(ref[1]).max())
        np.testing.assert_allclose(_cov(out[DD + D:2 * DD + D].reshape(D, D)), _cov(ref[2][0]),
                                   atol=1e-10 * np.abs(_cov(ref[2])).max())
        np.testing.assert_allclose(out[2 * DD + D:2 * DD + 2 * D], ref[3][0], atol=1e-10 * np.abs(ref[3]).max())
        np.testing.assert_allclose(_cov(out[2 * DD + 2 * D:].reshape(D, D)), _cov(ref[4][0]),
                                   atol=1e-10 * np.abs(_cov(ref[4])).max())


@pytest.mark.parametrize("name,kw,N,q,L,ks_max", [
    ("fitzhughnagumo", {}, 100, 3, 4, 4), ("fitzhughnagumo", {}, 100, 3, 4, 8), ("fitzhughnagumo", {}, 100, 3, 7, 2),
    ("fitzhughnagumo", {}, 1024, 3, 16, 8), ("fitzhughnagumo", {}, 1024, 3, 4, 384), ("fitzhughnagumo", {}, 777, 3, 3, 5),
    ("logistic", {}, 64, 3, 2, 3), ("rigid_body", {}, 256, 3, 4, 6), ("lotkavolterra", {}, 300, 2, 5, 7),
])
def test_hybrid_tree_schedule_matches_oracle(lib, name, kw, N, q, L, ks_max):
    """The library's HYBRID tree schedule (up-sweep to the first narrow level, Kogge-Stone prefix scan over its nodes for
    the filter, the mirrored suffix scan in element form for the smoother, down-sweeps from that level) restated on the
    host with the device combines: same results as the reference's scans, for base levels of every shape (odd node
    counts, non-powers of two, the library's own KS_MAX = 384)."""
    ivp = getattr(ivps, name)(**kw)
    ts = np.linspace(ivp.t0, ivp.tmax, N)
    setup = O.set_up_solver(ivp, ts, q)
    st = O.get_initial_trajectory(setup)
    dom = O.linearize_at(setup, st.mean[1:])
    d = setup["d"]
    D = d * (q + 1)
    qL = np.ascontiguousarray(O.preconditioned_discretize_1d(q)[1])
    x0 = np.concatenate([setup["x0"].mean, setup["x0"].chol.ravel()])
    H, c = np.ascontiguousarray(dom.H), np.ascontiguousarray(dom.b)
    res = {}
    for tag, km in (("hybrid", ks_max), ("plain", 0)):
        means, chols = st.mean.copy(), np.zeros((N, D, D))
        fm, fc, sc = np.zeros((N, D)), np.zeros((N, D, D)), np.zeros(8)
        rc = lib.hs_linear_filtsmooth_hybrid(d, q, ctypes.c_long(N), ctypes.c_long(L), ctypes.c_long(km), _p(qL),
                                             _p(x0), _p(H), _p(c), _p(means), _p(chols), _p(fm), _p(fc), 0, _p(sc))
        assert rc == 0
        res[tag] = (means, chols, fm, fc, sc)
    means, chols, fm, fc, sc = res["hybrid"]
    filt, nll, _, ssq, ssqp = O.linear_noiseless_filtering(setup["x0"], setup["dtm"], dom)
    out, obj = O.smoothing(setup["dtm"], filt)
    assert np.abs(fm - filt.mean).max() <= 1e-9 * np.abs(filt.mean).max()
    assert np.abs(_cov(fc) - _cov(filt.chol)).max() <= 1e-10 * np.abs(_cov(filt.chol)).max()
    assert np.abs(means - out.mean).max() <= 1e-9 * np.abs(out.mean).max()
    assert np.abs(_cov(chols) - _cov(out.chol)).max() <= 1e-10 * np.abs(_cov(out.chol)).max()
    assert abs(sc[0] - nll) <= 1e-9 * abs(nll) and abs(sc[1] - obj) <= 1e-9 * abs(obj)
    assert abs(sc[3] - ssqp) <= 1e-9 * ssqp
    # and the two schedules agree with each other far below the gates (same leaves, different association order)
    assert np.abs(means - res["plain"][0]).max() <= 1e-10 * np.abs(out.mean).max()
