"""Large-state ("tile") kernels on the host: the CTA-cooperative device code of csrc/pof_tile.cuh (leaf recursions,
tree operators, chunk-level smoothing op) compiled for the CPU and driven with the chunking / tree schedule of the
CUDA library, checked against the oracle -- including noisy observations (cholR != 0) and the D = 64 Lorenz-96
problem of BASELINE config 5.

Race check: every piece of work in that code is a `Team::each(n, body)` followed by a CTA barrier, and nothing outside
a body touches memory.  The simulator executes the iterations of every `each` forward, backward and permuted; bitwise
identical results under all three orders mean no iteration depends on another one of the same `each`, i.e. the CUDA
kernels (one iteration per thread between two barriers) have no shared-memory races.  Shared memory is poisoned with
NaN before every chunk / node, so a read of a never-written entry would surface as well."""
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


def _p(a):
    return None if a is None else a.ctypes.data_as(P)


def _cov(L):
    return L @ np.swapaxes(L, -1, -2)


def _problem(name, kw, N, q, noisy):
    ivp = getattr(ivps, name)(**kw)
    ts = np.linspace(ivp.t0, ivp.tmax, N)
    setup = O.set_up_solver(ivp, ts, q)
    st = O.get_initial_trajectory(setup)
    dom = O.linearize_at(setup, st.mean[1:])
    if noisy:
        rng = np.random.default_rng(1)
        d = setup["d"]
        R = np.tril(0.05 * rng.standard_normal((N - 1, d, d))) + 0.1 * np.eye(d)
        dom = O.AffineModel(dom.H, dom.b, R)
    return setup, st, dom


def _run(lib, setup, st, dom, q, L, order=0, noisy=False, compact=None, dense_model=None, reg=1):
    d = setup["d"]
    D = d * (q + 1)
    N = st.mean.shape[0]
    qL = np.ascontiguousarray(O.preconditioned_discretize_1d(q)[1])
    x0 = np.concatenate([setup["x0"].mean, setup["x0"].chol.ravel()])
    means, chols = st.mean.copy(), np.zeros((N, D, D))
    fm, fc, sc = np.zeros((N, D)), np.zeros((N, D, D)), np.zeros(8)
    H, c = np.ascontiguousarray(dom.H), np.ascontiguousarray(dom.b)
    R = np.ascontiguousarray(dom.cholR) if noisy else None
    Jc, s0, s1 = (None, 0.0, 0.0) if compact is None else compact
    rc = lib.hs_tile_linear_filtsmooth(
        d, q, ctypes.c_long(N), ctypes.c_long(L), _p(qL), _p(x0), None if Jc is not None else _p(H),
        None if Jc is not None else _p(c), _p(Jc), ctypes.c_double(s0), ctypes.c_double(s1), _p(R),
        None if dense_model is None else _p(dense_model[0]), None if dense_model is None else _p(dense_model[1]),
        _p(means), _p(chols), _p(fm), _p(fc), 0, _p(sc), order, reg)
    assert rc == 0
    return means, chols, fm, fc, sc


CASES = [
    ("fitzhughnagumo", {}, 100, 3, 7, False), ("fitzhughnagumo", {}, 100, 3, 200, False),
    ("logistic", {}, 64, 1, 5, False), ("rigid_body", {}, 256, 3, 8, False),
    ("henonheiles", {"tmax": 10.0}, 128, 2, 8, False), ("lotkavolterra", {}, 300, 2, 1, False),
    ("fitzhughnagumo", {}, 100, 3, 7, True), ("rigid_body", {}, 120, 2, 9, True),
    ("lorenz96", {"tmax": 1.0, "d": 8}, 60, 2, 7, True),
    ("lorenz96", {"tmax": 1.0}, 40, 3, 6, False),  # d = 16, q = 3: D = 64 (BASELINE config 5)
]


@pytest.mark.parametrize("name,kw,N,q,L,noisy", CASES)
def test_tile_pass_matches_oracle_and_is_order_independent(lib, name, kw, N, q, L, noisy):
    setup, st, dom = _problem(name, kw, N, q, noisy)
    means, chols, fm, fc, sc = _run(lib, setup, st, dom, q, L, 0, noisy)
    filt, nll, _, ssq, ssqp = O.linear_noiseless_filtering(setup["x0"], setup["dtm"], dom)
    out, obj = O.smoothing(setup["dtm"], filt)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(fm, filt.mean) <= 1e-9 and rel(means, out.mean) <= 1e-9
    assert rel(_cov(fc), _cov(filt.chol)) <= 1e-10 and rel(_cov(chols), _cov(out.chol)) <= 1e-10
    assert abs(sc[0] - nll) <= 1e-9 * abs(nll) and abs(sc[1] - obj) <= 1e-9 * abs(obj)
    assert abs(sc[3] - ssqp) <= 1e-9 * ssqp
    if not noisy:  # the reference's sigma^2 formula depends on QR sign conventions (utils.py:110-112)
        assert abs(sc[2] - ssq) <= 1e-2 * ssq
    moved = np.sum(~(np.abs(st.mean - out.mean) <= 1e-8 + 1e-13 * np.abs(out.mean)))
    assert abs(sc[4] - moved) <= 2
    for order in (1, 2):
        again = _run(lib, setup, st, dom, q, L, order, noisy)
        for a, b in zip((means, chols, fm, fc, sc), again):
            assert np.array_equal(a, b, equal_nan=True), f"iteration order {order} changes the result: race"
    # the default configuration of the kernels: shared-memory Householder sweeps everywhere (the run above used the
    # register-resident sweeps); same results up to rounding, and order-independent as well
    smem = [_run(lib, setup, st, dom, q, L, order, noisy, reg=0) for order in (0, 1)]
    for a, b in zip(smem[0], smem[1]):
        assert np.array_equal(a, b, equal_nan=True)
    assert rel(smem[0][0], means) <= 1e-10 and rel(_cov(smem[0][1]), _cov(chols)) <= 1e-10


def test_tile_lorenz96_compact_linearisation(lib):
    """k_linearize_l96's body against the oracle's symbolic Jacobian, and the compact [J_f | c] form the fused
    iteration feeds to the tile kernels against the dense (H, c) form"""
    q, N, L = 3, 30, 5
    setup, st, dom = _problem("lorenz96", {"tmax": 0.5}, N, q, False)
    d, D, n = 16, 64, N - 1
    s0, s1 = setup["E0"][0, 0], setup["E1"][0, 1]
    H, c, Jc = np.zeros((n, d, D)), np.zeros((n, d)), np.zeros((n, d * d + d))
    m1 = np.ascontiguousarray(st.mean[1:])
    assert lib.hs_linearize_l96(ctypes.c_double(8.0), ctypes.c_long(n), d, q, ctypes.c_double(s0), ctypes.c_double(s1),
                                _p(m1), _p(H), _p(c), _p(Jc)) == 0
    np.testing.assert_allclose(H, dom.H, rtol=0, atol=1e-13 * np.abs(dom.H).max())
    np.testing.assert_allclose(c, dom.b, rtol=0, atol=1e-12 * np.abs(dom.b).max())
    dense = _run(lib, setup, st, O.AffineModel(H, c, dom.cholR), q, L)
    compact = _run(lib, setup, st, dom, q, L, compact=(Jc, s0, s1))
    for a, b in zip(dense, compact):
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-12 * max(1.0, np.abs(a).max()))


@pytest.mark.parametrize("D", [6, 24, 64])
def test_tile_tree_operators_match_oracle(lib, D):
    rng = np.random.default_rng(D)
    pack = lambda e: np.concatenate([x.reshape(-1) for x in e])
    DD = D * D
    for order in (0, 1, 2):
        e1, e2 = [(rng.standard_normal((1, D, D)) / np.sqrt(D), rng.standard_normal((1, D)),
                   np.tril(rng.standard_normal((1, D, D))) / np.sqrt(D), rng.standard_normal((1, D)),
                   np.tril(rng.standard_normal((1, D, D))) / np.sqrt(D)) for _ in range(2)]
        out = np.zeros(3 * DD + 2 * D)
        assert lib.hs_tile_filter_combine(D, _p(pack(e1)), _p(pack(e2)), _p(out), 0, order) == 0
        ref = O.sqrt_filtering_operator(e1, e2)
        tol = lambda x: 1e-9 * np.abs(x).max()
        np.testing.assert_allclose(out[:DD].reshape(D, D), ref[0][0], rtol=0, atol=tol(ref[0]))
        np.testing.assert_allclose(out[DD:DD + D], ref[1][0], rtol=0, atol=tol(ref[1]))
        np.testing.assert_allclose(_cov(out[DD + D:2 * DD + D].reshape(D, D)), _cov(ref[2][0]), rtol=0,
                                   atol=tol(_cov(ref[2])))
        np.testing.assert_allclose(out[2 * DD + D:2 * DD + 2 * D], ref[3][0], rtol=0, atol=tol(ref[3]))
        np.testing.assert_allclose(_cov(out[2 * DD + 2 * D:].reshape(D, D)), _cov(ref[4][0]), rtol=0,
                                   atol=tol(_cov(ref[4])))
        s1, s2 = [(rng.standard_normal((1, D)), rng.standard_normal((1, D, D)) / np.sqrt(D),
                   np.tril(rng.standard_normal((1, D, D)))) for _ in range(2)]
        out = np.zeros(2 * DD + D)
        assert lib.hs_tile_smooth_combine(D, _p(pack(s1)), _p(pack(s2)), _p(out), 0, order) == 0
        ref = O.sqrt_smoothing_operator(s1, s2)
        np.testing.assert_allclose(out[:D], ref[0][0], rtol=0, atol=tol(ref[0]))
        np.testing.assert_allclose(out[D:D + DD].reshape(D, D), ref[1][0], rtol=0, atol=tol(ref[1]))
        np.testing.assert_allclose(_cov(out[D + DD:].reshape(D, D)), _cov(ref[2][0]), rtol=0, atol=tol(_cov(ref[2])))


def test_tile_shared_memory_fits_config5(lib):
    """d = 16, q = 3 (D = 64): every tile kernel's dynamic shared memory fits the 227 KB a B200 CTA can have"""
    for which in range(4):
        assert lib.hs_tile_smem_bytes(64, 16, which) <= 227 * 1024


@pytest.mark.parametrize("name,q,N,L,noisy", [("fitzhughnagumo", 2, 60, 7, False), ("rigid_body", 3, 45, 4, True)])
def test_tile_general_transition_models_nonuniform_grid(lib, name, q, N, L, noisy):
    """per-step dense (F_k, QL_k): the reference's non-preconditioned models on a NON-uniform grid
    (pof/transitions.py:71-99 `non_preconditioned_discretize`, pof/convenience.py:48-73 `set_up_solver_no_precond`)"""
    ivp = getattr(ivps, name)()
    d = ivp.y0.shape[0]
    D = d * (q + 1)
    ts = ivp.t0 + (ivp.tmax - ivp.t0) * 0.3 * np.linspace(0, 1, N) ** 1.5
    F0, QL0 = O.preconditioned_discretize(d, q)
    Fs, QLs = np.empty((N - 1, D, D)), np.empty((N - 1, D, D))
    for k, dt in enumerate(np.diff(ts)):
        Pk, PIk = O.nordsieck_preconditioner(d, q, dt)
        Fs[k], QLs[k] = Pk @ F0 @ PIk, Pk @ QL0
    E0, E1 = O.projection_matrix(d, q, 0), O.projection_matrix(d, q, 1)
    x0 = O.taylor_mode_init(ivp, q)
    st = O.constant_init(ivp, q, N)
    setup = dict(ivp=ivp, ts=ts, dtm=O.TransitionModel(Fs, QLs), x0=x0, E0=E0, E1=E1, order=q, d=d)
    dom = O.linearize_at(setup, st.mean[1:])
    if noisy:
        rng = np.random.default_rng(2)
        dom = O.AffineModel(dom.H, dom.b, np.tril(0.05 * rng.standard_normal((N - 1, d, d))) + 0.1 * np.eye(d))
    res = [_run(lib, setup, st, dom, q, L, order, noisy, dense_model=(Fs, QLs)) for order in (0, 1, 2)]
    means, chols, fm, fc, sc = res[0]
    filt, nll, _, ssq, ssqp = O.linear_noiseless_filtering(x0, setup["dtm"], dom)
    out, obj = O.smoothing(setup["dtm"], filt)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(fm, filt.mean) <= 1e-9 and rel(means, out.mean) <= 1e-9
    assert rel(_cov(fc), _cov(filt.chol)) <= 1e-9 and rel(_cov(chols), _cov(out.chol)) <= 1e-9
    assert abs(sc[0] - nll) <= 1e-9 * abs(nll) and abs(sc[1] - obj) <= 1e-9 * abs(obj)
    assert abs(sc[3] - ssqp) <= 1e-9 * ssqp
    for other in res[1:]:
        for a, b in zip(res[0], other):
            assert np.array_equal(a, b, equal_nan=True)


IVP_IDS = dict(logistic=0, lotkavolterra=1, vanderpol=2, fitzhughnagumo=3, rober=4, rigid_body=5, seir=6, threebody=7,
               henonheiles=8, lorenz96=9)


@pytest.mark.parametrize("name,kw,params,N,q", [
    ("fitzhughnagumo", {}, (0.7, 0.8, 1 / 12.5, 0.5), 50, 3), ("logistic", {}, (), 40, 2),
    ("rigid_body", {}, (-2.0, 1.25, -0.5), 60, 3), ("lorenz96", {"tmax": 0.5}, (8.0,), 24, 3),
])
def test_tile_sequential_eks_matches_oracle(lib, name, kw, params, N, q):
    """one-CTA sequential EKS (relinearised at the predicted mean inside the kernel) incl. d = 16, D = 64"""
    ivp = getattr(ivps, name)(**kw)
    ts = np.linspace(ivp.t0, ivp.tmax, N)
    setup = O.set_up_solver(ivp, ts, q)
    d = setup["d"]
    D = d * (q + 1)
    qL = np.ascontiguousarray(O.preconditioned_discretize_1d(q)[1])
    x0 = np.concatenate([setup["x0"].mean, setup["x0"].chol.ravel()])
    s0, s1 = setup["E0"][0, 0], setup["E1"][0, 1]
    p8 = np.zeros(8)
    p8[: len(params)] = params
    out, ell, obj, ssq = O.sequential_eks(setup)
    def run(order, reg):
        means, chols, sums = np.zeros((N, D)), np.zeros((N, D, D)), np.zeros(8)
        assert lib.hs_tile_seq_eks(d, q, ctypes.c_long(N), _p(qL), ctypes.c_double(s0), ctypes.c_double(s1),
                                   IVP_IDS[name], _p(p8), _p(x0), _p(means), _p(chols), _p(sums), order, reg) == 0
        return means, chols, sums[:4].copy()

    res = [run(order, 1) for order in (0, 1, 2)]  # register-resident sweeps
    means, chols, sums = res[0]
    m0, c0, _ = run(0, 0)  # shared-memory sweeps (the kernels' default)
    assert np.abs(m0 - means).max() <= 1e-10 * np.abs(means).max()
    assert np.abs(_cov(c0) - _cov(chols)).max() <= 1e-10 * np.abs(_cov(chols)).max()
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(means, out.mean) <= 1e-9
    assert rel(_cov(chols), _cov(out.chol)) <= 1e-9
    assert abs(-sums[0] - ell) <= 1e-9 * abs(ell) and abs(sums[3] - obj) <= 1e-9 * abs(obj)
    # the reference's sigma^2 (whiten solves with L^T, utils.py:110-112) depends on which of the valid innovation
    # factors the QR returns; the dependence grows with d (3 % at d = 16), cf. DESIGN.md section 4
    assert abs(sums[1] / (N - 1) / d - ssq) <= (1e-2 if d <= 4 else 1e-1) * ssq
    for other in res[1:]:
        for a, b in zip(res[0], other):
            assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("R,C,npiv,c0", [
    (32, 32, 8, -1), (24, 24, 24, -1), (80, 64, 16, -1), (10, 10, 2, -1),          # plain: one thread per row, shifting
    (48, 48, 24, 24), (24, 48, 24, 24), (36, 60, 18, 18),                          # pentagonal, 2 threads per row
    (128, 128, 64, 64), (64, 128, 64, 64), (64, 80, 64, 64), (24, 32, 24, 24),     # D = 64: scan (H=2), fold/smooth (H=4)
])
def test_register_sweeps_equal_shared_memory_sweep(lib, R, C, npiv, c0):
    """tile_tria_regp / tile_tria_reg1 (rows in registers, pivot row broadcast) against tile_tria_smem on random
    arrays, under all iteration orders (each_group's stage 2 runs the parts of a row forward / backward)"""
    def run(use_reg, order):
        rng = np.random.default_rng(7)
        ld = C + 1
        M = rng.standard_normal((R, ld))
        if c0 >= 0:
            for r in range(min(R, c0)):
                M[r, r + 1:c0] = 0.0
        assert lib.hs_tile_tria(_p(M), R, C, ld, npiv, c0, use_reg, order) == 0
        out = M[:, :C].copy()
        for r in range(min(R, npiv)):
            out[r, r + 1:] = 0.0  # eliminated entries of the pivot rows are dead
        return out

    ref = run(0, 0)
    res = [run(1, o) for o in (0, 1, 2)]
    assert np.abs(res[0] - ref).max() <= 1e-13 * np.abs(ref).max()
    for x in res[1:]:
        assert np.array_equal(res[0], x)
