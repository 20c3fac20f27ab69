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
