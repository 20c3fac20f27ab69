"""Host-side pieces of the facade that need no GPU: transition-model discretisation (reference pof/transitions.py,
tests/test_transtions.py), the Lorenz-96 problem (BASELINE config 5) against the oracle's symbolic definition, and
the Taylor-mode initial state (replacement of tornadox.init.TaylorMode) at d = 16."""
import math

import numpy as np
import pytest
import torch

from oracle import ivps as oivps
from oracle import pof_oracle as O


@pytest.mark.parametrize("dim", [1, 3, 5])
@pytest.mark.parametrize("order", [1, 3, 5])
def test_projection_and_discretize_shapes(native_lib, dim, order):
    """reference tests/test_transtions.py:14-47"""
    from pof.transitions import IWP, TransitionModel, get_transition_model, projection_matrix

    iwp = IWP(num_derivatives=order, wiener_process_dimension=dim)
    D = dim * (order + 1)
    E0, E1 = projection_matrix(iwp, 0), projection_matrix(iwp, 1)
    assert E0.shape == (dim, D) and E1.shape == (dim, D)
    x = np.random.default_rng(0).uniform(size=D)
    assert all(E0 @ x == x[0:: order + 1]) and all(E1 @ x == x[1:: order + 1])
    tm = get_transition_model(iwp, 0.1)
    assert isinstance(tm, TransitionModel) and tm.F.shape == (D, D) and tm.QL.shape == (D, D)


def test_non_preconditioned_model_is_the_iwp_discretisation(native_lib):
    """F(dt)[i, j] = dt^(j-i) / (j-i)! per block, and QL QL^T is the integrated-Wiener-process covariance
    (reference transitions.py:71-77 composed of :37-68)"""
    from pof.transitions import IWP, discretize_transitions, get_transition_model

    d, q, dt = 2, 3, 0.37
    iwp = IWP(num_derivatives=q, wiener_process_dimension=d)
    F, QL = get_transition_model(iwp, dt)
    F1 = np.array([[dt ** (j - i) / math.factorial(j - i) if j >= i else 0.0 for j in range(q + 1)]
                   for i in range(q + 1)])
    np.testing.assert_allclose(F, np.kron(np.eye(d), F1), rtol=1e-13, atol=1e-15)
    Q1 = np.array([[dt ** (2 * q + 1 - i - j) / ((2 * q + 1 - i - j) * math.factorial(q - i) * math.factorial(q - j))
                    for j in range(q + 1)] for i in range(q + 1)])
    np.testing.assert_allclose(QL @ QL.T, np.kron(np.eye(d), Q1), rtol=1e-11, atol=1e-18)
    assert np.abs(np.triu(QL, 1)).max() == 0.0
    ts = np.array([0.0, 0.1, 0.35, 0.4, 1.0])
    tm = discretize_transitions(iwp, times=ts)
    assert tuple(tm.F.shape) == (4, 8, 8) and tm.F.dtype == torch.float64
    for k, h in enumerate(np.diff(ts)):
        Fk, QLk = get_transition_model(iwp, h)
        np.testing.assert_allclose(tm.F[k].numpy(), Fk, rtol=1e-14, atol=0)
        np.testing.assert_allclose(tm.QL[k].numpy(), QLk, rtol=1e-14, atol=0)
        Pk, PIk = O.nordsieck_preconditioner(d, q, h)
        F0, QL0 = O.preconditioned_discretize(d, q)
        np.testing.assert_allclose(Fk, Pk @ F0 @ PIk, rtol=1e-13, atol=1e-18)


def test_lorenz96_matches_the_oracle_definition(native_lib):
    import pof.ivp
    from pof import initialization as init

    ivp, oivp = pof.ivp.lorenz96(tmax=1.0), oivps.lorenz96(tmax=1.0)
    assert ivp.dimension == 16 and ivp.t_span == (0.0, 1.0)
    np.testing.assert_allclose(ivp.y0.numpy(), oivp.y0, rtol=0, atol=0)
    y = torch.linspace(-2.0, 9.0, 16, dtype=torch.float64) ** 2 / 7.0
    np.testing.assert_allclose(ivp.f(None, y).numpy(), oivp.f(None, y.numpy()), rtol=1e-15, atol=1e-14)
    assert ivp.f._pof_builtin == (native_lib.IVP_IDS["lorenz96"], (8.0,))
    # Jacobian through autodiff (the path user-supplied vector fields take) against the symbolic one
    J = torch.func.jacfwd(lambda v: ivp.f(None, v))(y).numpy()
    np.testing.assert_allclose(J, oivp.jac(y.numpy()), rtol=1e-14, atol=1e-13)
    # Taylor-mode initial state, order 3: derivatives y^(k)(t0) by nested jvps vs exact symbolic differentiation
    x0 = init.taylor_mode_init(ivp.f, ivp.y0, 3)
    ox0 = O.taylor_mode_init(oivp, 3)
    np.testing.assert_allclose(x0.mean.numpy(), ox0.mean, rtol=1e-12, atol=1e-11)
    assert float(x0.chol.abs().max()) == 0.0
    small = pof.ivp.lorenz96(d=8)
    assert small.dimension == 8


def test_observation_model_variants(native_lib):
    """reference observations.py:35-83 (`linearize`, `linearize_ek0`, `uncertain_linearize`, `linearize_regularized`):
    shapes, the affine identity H m + b == f(m) (reference tests/test_observations.py:18-33), and the stacked
    Levenberg-Marquardt model [EK1 ; x ~ N(m, I / l)] with the reference's offset convention"""
    from pof.observations import (AffineModel, linearize, linearize_ek0, linearize_regularized, uncertain_linearize)
    from pof.utils import MVNSqrt

    d, q = 2, 2
    D = d * (q + 1)
    unit = lambda i: torch.eye(q + 1, dtype=torch.float64)[i:i + 1]
    E0 = torch.kron(torch.eye(d, dtype=torch.float64), unit(0))
    E1 = torch.kron(torch.eye(d, dtype=torch.float64), unit(1))
    f = lambda x: E1 @ x - torch.sin(E0 @ x)
    m = torch.linspace(0.1, 0.6, D, dtype=torch.float64)
    L = torch.tril(torch.full((D, D), 0.1, dtype=torch.float64)) + 0.5 * torch.eye(D, dtype=torch.float64)
    x = MVNSqrt(m, L)
    lin = linearize(f, x)
    assert isinstance(lin, AffineModel) and lin.H.shape == (d, D) and lin.b.shape == (d,)
    assert float(lin.cholR.abs().max()) == 0.0
    np.testing.assert_allclose((lin.H @ m + lin.b).numpy(), f(m).numpy(), rtol=1e-14)
    ek0 = linearize_ek0(f, x)
    np.testing.assert_allclose(ek0.H.numpy(), E1.numpy())
    np.testing.assert_allclose((ek0.H @ m + ek0.b).numpy(), f(m).numpy(), rtol=1e-14)
    unc = uncertain_linearize(f, x)
    np.testing.assert_allclose((unc.cholR @ unc.cholR.T).numpy(), (lin.H @ L @ L.T @ lin.H.T).numpy(), rtol=1e-12)
    reg = linearize_regularized(f, x, 4.0)
    assert reg.H.shape == (d + D, D) and reg.b.shape == (d + D,) and reg.cholR.shape == (d + D, d + D)
    np.testing.assert_allclose(reg.H[:d].numpy(), lin.H.numpy())
    np.testing.assert_allclose(reg.H[d:].numpy(), np.eye(D))
    np.testing.assert_allclose(reg.b.numpy(), np.concatenate([f(m).numpy(), -m.numpy()]))
    np.testing.assert_allclose(np.diag(reg.cholR.numpy()), np.concatenate([np.zeros(d), np.full(D, 0.5)]))


def test_iterator_and_batch_modules_import(native_lib):
    """the generator API and the batch solver are importable without a GPU (they only launch work when called)"""
    import pof.batch
    import pof.iterators

    assert callable(pof.iterators.ieks_iterator) and callable(pof.iterators.qpm_ieks_iterator)
    assert callable(pof.batch.solve_batch) and callable(pof.iterators.lm_ieks_iterator)


def test_oracle_qpm_reaches_the_plain_ieks_solution():
    """oracle restatement of the quadratic-penalty iterator (iterators.py:53-112): it ends with reg = 0 at the fixed
    point of the plain IEKS"""
    oivp = oivps.logistic()
    ts = np.linspace(0, 10, 40)
    s = O.set_up_solver(oivp, ts, 2)
    regs = []
    for st, nll, obj, reg in O.qpm_ieks_iterator(s, O.get_initial_trajectory(s)):
        regs.append(reg)
        assert len(regs) < 400
    assert regs[0] == 1e20 and regs[-1] == 0.0 and all(a >= b for a, b in zip(regs, regs[1:]))
    ys, info = O.solve(oivp, ts, 2)
    np.testing.assert_allclose(st.mean @ s["E0"].T, ys.mean, rtol=0, atol=1e-10)


def test_oracle_prior_init_is_the_taylor_prediction():
    """oracle restatement of init="prior" (initialization.py:66-89): row k is the Taylor polynomial of the initial
    derivatives over the ABSOLUTE time ts[k] (quirk Q6), factor -P_k QL under LAPACK's sign convention"""
    oivp = oivps.logistic()
    q = 3
    ts = np.array([0.0, 0.5, 1.0, 2.5])
    st = O.prior_init(oivp, q, ts)
    m0 = O.taylor_mode_init(oivp, q).mean
    for k, t in enumerate(ts[1:], start=1):
        want = [sum(t ** (j - i) / math.factorial(j - i) * m0[j] for j in range(i, q + 1)) for i in range(q + 1)]
        np.testing.assert_allclose(st.mean[k], want, rtol=1e-12)
        P, _ = O.nordsieck_preconditioner(1, q, t)
        _, QL = O.preconditioned_discretize(1, q)
        np.testing.assert_allclose(st.chol[k], -P @ QL, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(st.mean[0], m0)


def test_small_transition_helpers(native_lib):
    """reference transitions.py:28-34, 54-61, 81-82 and step.py:26-30"""
    from pof.step import inflate
    from pof.transitions import (IWP, hilbert, nordsieck_preconditioner, nordsieck_preconditioner_1d, pascal,
                                 preconditioned_discretize_1d, projection_matrix, projection_matrix_1d)
    from pof.utils import MVNSqrt

    iwp = IWP(num_derivatives=3, wiener_process_dimension=2)
    A, L = preconditioned_discretize_1d(iwp)
    np.testing.assert_allclose(A, np.flip(pascal(4)))
    np.testing.assert_allclose(L @ L.T, np.flip(hilbert(4)), rtol=1e-12)
    P1, PI1 = nordsieck_preconditioner_1d(iwp, 0.3)
    P, PI = nordsieck_preconditioner(iwp, 0.3)
    np.testing.assert_allclose(np.kron(np.eye(2), P1), P)
    np.testing.assert_allclose(P1 @ PI1, np.eye(4), rtol=1e-13)
    np.testing.assert_allclose(np.kron(np.eye(2), projection_matrix_1d(iwp, 1)), projection_matrix(iwp, 1))
    x = inflate(MVNSqrt(torch.zeros(3, dtype=torch.float64), torch.zeros((3, 3), dtype=torch.float64)))
    np.testing.assert_allclose(x.chol.numpy(), 1e-3 * np.eye(3))
