"""IWP prior in preconditioned coordinates (reference pof/transitions.py).

Host-side, one-off: the discretisation does not depend on dt (transitions.py:37-50), so F and QL are computed once
and kept un-replicated; the kernels take the (q+1)x(q+1) block qL and know F's Pascal structure at compile time.
"""
from typing import NamedTuple

import numpy as np
import scipy.linalg
import scipy.special
import torch


class TransitionModel(NamedTuple):
    """Linear transition model in square-root form: x' | x ~ N(Fx, QL*QLt) (reference transitions.py:15-19)."""

    F: torch.Tensor
    QL: torch.Tensor


class IWP(NamedTuple):
    wiener_process_dimension: int
    num_derivatives: int


def hilbert(n):
    """reference transitions.py:28-29"""
    return scipy.linalg.hilbert(n)


def pascal(n):
    """reference transitions.py:32-34 (lower-triangular Pascal matrix)"""
    return scipy.linalg.pascal(n, kind="lower", exact=False)


def nordsieck_preconditioner_1d(iwp: IWP, dt):
    """reference transitions.py:54-61"""
    sv, svi = nordsieck_scalings(iwp, dt)
    return np.diag(sv), np.diag(svi)


def projection_matrix_1d(iwp: IWP, derivative_to_project_onto):
    """reference transitions.py:81-82"""
    return np.eye(1, iwp.num_derivatives + 1, derivative_to_project_onto)


def preconditioned_discretize_1d(iwp: IWP):
    """reference transitions.py:37-41"""
    q = iwp.num_derivatives
    A_1d = np.flip(scipy.linalg.pascal(q + 1, kind="lower", exact=False))
    Q_1d = np.flip(scipy.linalg.hilbert(q + 1))
    return np.ascontiguousarray(A_1d), np.linalg.cholesky(Q_1d)


def preconditioned_discretize(iwp: IWP):
    """reference transitions.py:44-50"""
    A_1d, L_Q1d = preconditioned_discretize_1d(iwp)
    Id = np.eye(iwp.wiener_process_dimension)
    return np.kron(Id, A_1d), np.kron(Id, L_Q1d)


def nordsieck_scalings(iwp: IWP, dt):
    q = iwp.num_derivatives
    powers = np.arange(q, -1, -1)
    scales = scipy.special.factorial(powers)
    powers = powers + 0.5
    return (np.abs(dt) ** powers) / scales, (np.abs(dt) ** (-powers)) * scales


def nordsieck_preconditioner(iwp: IWP, dt):
    """reference transitions.py:53-68"""
    sv, svi = nordsieck_scalings(iwp, dt)
    Id = np.eye(iwp.wiener_process_dimension)
    return np.kron(Id, np.diag(sv)), np.kron(Id, np.diag(svi))


def projection_matrix(iwp: IWP, derivative_to_project_onto):
    """reference transitions.py:80-88"""
    return np.kron(np.eye(iwp.wiener_process_dimension), np.eye(1, iwp.num_derivatives + 1, derivative_to_project_onto))


def non_preconditioned_discretize(iwp: IWP, dt):
    """reference transitions.py:71-77: (P F PI, P QL) for one step size"""
    P, PI = nordsieck_preconditioner(iwp, dt)
    F, QL = preconditioned_discretize(iwp)
    return P @ F @ PI, P @ QL


def get_transition_model(iwp: IWP, dt):
    """reference transitions.py:91-93 (numpy arrays; `discretize_transitions` stacks them on the device)"""
    return TransitionModel(*non_preconditioned_discretize(iwp, dt))


def discretize_transitions(iwp: IWP, times=None, steps=None, device=None):
    """reference transitions.py:96-100: per-step models for arbitrary (non-uniform) grids,
    TransitionModel(F (n,D,D), QL (n,D,D)) -- the input of the general pass `pof_linear_filtsmooth_general_f64`."""
    if steps is None:
        times = np.asarray(times, dtype=np.float64)
        steps = times[1:] - times[:-1]
    F0, QL0 = preconditioned_discretize(iwp)
    sv = np.stack([nordsieck_scalings(iwp, dt)[0] for dt in steps])    # (n, q+1)
    svi = np.stack([nordsieck_scalings(iwp, dt)[1] for dt in steps])
    d = iwp.wiener_process_dimension
    Pd, PId = np.tile(sv, (1, d)), np.tile(svi, (1, d))                  # diagonals of P_k, PI_k: (n, D)
    Fs = Pd[:, :, None] * F0[None] * PId[:, None, :]
    QLs = Pd[:, :, None] * QL0[None]
    tt = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=device)
    return TransitionModel(tt(Fs), tt(QLs))
