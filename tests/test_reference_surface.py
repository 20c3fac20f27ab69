"""The reference's own tests, restated against this package's API (same imports, same calls): tests/test_filtsmooth.py
(sequential EKF / RTS vs parallel filter / smoother on a 1-d Wiener process), tests/test_initialization.py (`_prior_init`,
`updated_prior_init`), plus the small public helpers of pof/utils.py against the oracle.  These entry points are
device-agnostic torch code (baseline / cross-check paths and shapes the CUDA kernels do not take), so they run here on
CPU tensors; the CUDA paths have their own `-m gpu` tests."""
import warnings

import numpy as np
import pytest
import torch

from oracle import ivps as oivps
from oracle import pof_oracle as O

T = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64)


def test_equality_like_the_reference(native_lib):
    """reference tests/test_filtsmooth.py:16-63, line by line"""
    from pof.convenience import discretize_transitions, linearize_observation_model
    from pof.observations import NonlinearModel
    from pof.parallel_filtsmooth import linear_noiseless_filtering as pfilt
    from pof.parallel_filtsmooth import smoothing as psmooth
    from pof.sequential_filtsmooth import extended_kalman_filter as sfilt
    from pof.sequential_filtsmooth import smoothing as ssmooth
    from pof.transitions import IWP, projection_matrix
    from pof.utils import MVNSqrt

    iwp = IWP(num_derivatives=0, wiener_process_dimension=1)
    E0 = T(projection_matrix(iwp, 0))
    obsmod = NonlinearModel(lambda x: E0 @ x)
    x0 = MVNSqrt(torch.zeros(1, dtype=torch.float64), torch.zeros((1, 1), dtype=torch.float64))
    disc_transmod = discretize_transitions(iwp, np.arange(10))

    out_ekf, _, _ = sfilt(x0, disc_transmod, obsmod)
    out_eks, _ = ssmooth(disc_transmod, out_ekf)
    assert out_ekf.mean.shape == out_eks.mean.shape and out_ekf.chol.shape == out_eks.chol.shape

    N = disc_transmod.F.shape[0]
    traj = MVNSqrt(x0.mean[None].repeat(N, 1), x0.chol[None].repeat(N, 1, 1))
    lin_obsmod = linearize_observation_model(obsmod, traj)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        out_pkf, _, _, _ = pfilt(x0, disc_transmod, lin_obsmod)
    assert any("torch library calls" in str(x.message) for x in w)  # q = 0 is outside the kernels: said, not silent
    out_pks, _ = psmooth(disc_transmod, out_pkf)
    assert out_pkf.mean.shape == out_pks.mean.shape and out_pkf.chol.shape == out_pks.chol.shape
    assert out_ekf.mean.shape == out_pkf.mean.shape and out_ekf.chol.shape == out_pkf.chol.shape
    assert bool((out_ekf.mean == out_pkf.mean).all()) and bool((out_ekf.chol == out_pkf.chol).all())


@pytest.mark.parametrize("name,N,q", [("logistic", 30, 2), ("lotkavolterra", 25, 1), ("fitzhughnagumo", 40, 3)])
def test_sequential_and_parallel_forms_agree_with_the_oracle(native_lib, name, N, q):
    """non-degenerate version: per-step models on a grid, affine observations from a trajectory; the sequential loops,
    the stand-alone parallel filter / smoother and the oracle give the same marginals"""
    from pof.observations import AffineModel
    from pof.parallel_filtsmooth.library_pass import linear_noiseless_filtering_library
    from pof.parallel_filtsmooth import smoothing as psmooth
    from pof.sequential_filtsmooth import linear_noiseless_filter, smoothing as ssmooth
    from pof.transitions import TransitionModel
    from pof.utils import MVNSqrt

    ivp = getattr(oivps, name)()
    ts = np.linspace(0.0, 5.0, N)
    setup = O.set_up_solver(ivp, ts, q)
    st = O.ieks_step(setup, O.get_initial_trajectory(setup))[0]
    dom = O.linearize_at(setup, st.mean[1:])
    x0, dtm = setup["x0"], setup["dtm"]
    filt_o, nll_o, _, ssq_o, _ = O.linear_noiseless_filtering(x0, dtm, dom)
    out_o, obj_o = O.smoothing(dtm, filt_o)
    tx0 = MVNSqrt(T(x0.mean), T(x0.chol))
    tdtm = TransitionModel(T(dtm.F), T(dtm.QL))
    tdom = AffineModel(T(dom.H), T(dom.b), T(dom.cholR))
    cov = lambda L: L @ L.transpose(-1, -2)
    fs, ell, ssq_s = linear_noiseless_filter(tx0, tdtm, tdom)
    ss, obj_s = ssmooth(tdtm, fs)
    fp, nll_p, _, _ = linear_noiseless_filtering_library(tx0, tdtm, tdom)
    sp, obj_p = psmooth(tdtm, fp)
    Co = T(out_o.chol @ np.swapaxes(out_o.chol, -1, -2))
    for out, obj in ((ss, obj_s), (sp, obj_p)):
        # (sequential recursion vs the oracle's odd/even scan: different association orders)
        scale = np.abs(out_o.mean).max(axis=0)
        assert (np.abs(out.mean.numpy() - out_o.mean) <= 1e-7 * scale + 1e-12).all()
        assert float((cov(out.chol) - Co).abs().max()) <= 1e-7 * float(Co.abs().max())
        assert abs(float(obj) - obj_o) <= 1e-8 * abs(obj_o) + 1e-12
    assert abs(float(ell) + nll_o) <= 1e-9 * abs(nll_o) + 1e-9  # sequential: +sum loglik (quirk Q7)
    assert abs(float(nll_p) - nll_o) <= 1e-9 * abs(nll_o) + 1e-9
    assert np.isfinite(float(ssq_s)) and float(ssq_s) > 0


def test_prior_inits_like_the_reference(native_lib):
    """reference tests/test_initialization.py:32-58: `_prior_init` (row k = x0 through transition k alone, quirk Q6)
    against the oracle restatement, `updated_prior_init` shapes and consistency with its observation"""
    from pof.initialization import _prior_init, updated_prior_init
    from pof.observations import NonlinearModel
    from pof.transitions import TransitionModel
    from pof.utils import MVNSqrt

    ivp = oivps.logistic()
    order, ts = 3, np.arange(0.0, 2.0 + 0.25, 0.25)
    traj_o = O.prior_init(ivp, order, ts)
    setup = O.set_up_solver(ivp, ts, order)
    x0_raw = O.taylor_mode_init(ivp, order)  # non-preconditioned coordinates, as `prior_init` uses them
    from pof.transitions import IWP, discretize_transitions

    dtm = discretize_transitions(IWP(num_derivatives=order, wiener_process_dimension=1), steps=ts[1:])
    x0 = MVNSqrt(T(x0_raw.mean), T(x0_raw.chol))
    traj = _prior_init(x0=x0, dtm=dtm)
    assert traj.mean.shape == (len(ts), order + 1) and traj.chol.shape == (len(ts), order + 1, order + 1)
    ok = np.isfinite(traj_o.mean).all(axis=1)
    np.testing.assert_allclose(traj.mean.numpy()[ok], traj_o.mean[ok], rtol=1e-12, atol=1e-14)
    C = (traj.chol @ traj.chol.transpose(-1, -2)).numpy()
    Co = traj_o.chol @ np.swapaxes(traj_o.chol, -1, -2)
    np.testing.assert_allclose(C[ok], Co[ok], rtol=1e-10, atol=1e-14)
    E0, E1 = T(np.eye(order + 1)[0:1]), T(np.eye(order + 1)[1:2])
    om = NonlinearModel(lambda x: E1 @ x - (E0 @ x) * (1 - E0 @ x))  # logistic vector field, y' = y (1 - y)
    upd = updated_prior_init(x0=x0, dtm=dtm, om=om)
    assert upd.mean.shape == traj.mean.shape and upd.chol.shape == traj.chol.shape
    # row 0 is x0 (zero covariance) updated on a noiseless observation: 0 / 0 upstream as here; the predicted rows are finite
    assert bool(torch.isfinite(upd.mean[1:]).all()) and bool(torch.isfinite(upd.chol[1:]).all())
    res = torch.stack([om(m) for m in upd.mean[1:]])
    lin = torch.stack([om(m) for m in traj.mean[1:]])
    assert float(res.abs().max()) < float(lin.abs().max())  # the update moves every state towards f(x) = 0


def test_utils_helpers_match_the_oracle(native_lib):
    from pof.utils import append_zeros_along_new_axis, mvn_loglikelihood, objective_function_value, tria, whiten

    rng = np.random.default_rng(0)
    A = rng.standard_normal((4, 9))
    L = T(A)
    cov = lambda M: M @ M.T
    np.testing.assert_allclose(cov(tria(L).numpy()), cov(O.tria(A)), rtol=1e-12)
    assert float(torch.triu(tria(L), 1).abs().max()) == 0.0
    chol = np.tril(rng.standard_normal((4, 4))) + 3 * np.eye(4)
    x = rng.standard_normal(4)
    np.testing.assert_allclose(float(mvn_loglikelihood(T(x), T(chol))), O.mvn_loglikelihood(x, chol), rtol=1e-12)
    np.testing.assert_allclose(whiten(T(x), T(chol)).numpy(), O.whiten(x, chol), rtol=1e-12)
    F = rng.standard_normal((4, 4))
    m, mn = rng.standard_normal(4), rng.standard_normal(4)
    np.testing.assert_allclose(float(objective_function_value(T(mn), T(m), (T(F), T(chol)))),
                               O.objective_function_value(mn, m, F, chol), rtol=1e-12)
    z = append_zeros_along_new_axis(T(x), 3)
    assert z.shape == (4, 4) and bool((z[0] == T(x)).all()) and float(z[1:].abs().max()) == 0.0


def test_small_initialisers_like_the_reference(native_lib):
    """reference initialization.py:25-40 (`uncertain_init`), :58-72 (`classic_to_init`), :42-56 (`constant_init`)"""
    import pof.ivp
    from pof.initialization import classic_to_init, constant_init, uncertain_init

    ivp = pof.ivp.lotkavolterra()
    y0 = torch.as_tensor(np.asarray(ivp.y0), dtype=torch.float64)
    f0 = torch.as_tensor(np.asarray(ivp.f(None, y0)), dtype=torch.float64)
    x0 = uncertain_init(ivp.f, ivp.y0, 3, var=4.0)
    assert x0.mean.shape == (8,) and x0.chol.shape == (8, 8)
    np.testing.assert_allclose(x0.mean.numpy()[[0, 4]], y0.numpy())
    np.testing.assert_allclose(x0.mean.numpy()[[1, 5]], f0.numpy())
    np.testing.assert_allclose(np.diag(x0.chol.numpy()), [0, 0, 2, 2, 0, 0, 2, 2])
    ys = torch.stack([y0, 2 * y0, 3 * y0])
    tr = classic_to_init(ys=ys, order=2, f=ivp.f)
    assert tr.mean.shape == (3, 6) and tr.chol.shape == (3, 6, 6) and float(tr.chol.abs().max()) == 0.0
    np.testing.assert_allclose(tr.mean.numpy()[:, [0, 3]], ys.numpy())
    c = constant_init(y0=ivp.y0, order=2, ts=np.arange(5), f=ivp.f)
    np.testing.assert_allclose(c.mean.numpy(), np.repeat(tr.mean.numpy()[:1], 5, axis=0))
