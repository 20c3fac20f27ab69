"""The reference's sequential square-root filter / smoother as written there -- one step after the other
(reference pof/sequential_filtsmooth/filter.py:9-92, smoother.py:8-48).  Baseline and cross-check only (SURVEY 8a row
a21): O(N) tiny torch operations on the tensors' device, for ANY transition model, observation model and dimension
(the reference's own tests run them on a 1-dimensional Wiener process, q = 0).  The fast sequential paths are
`linear_filtsmooth` (the CUDA pass with one chunk) and `pof_sequential_eks_f64` (one GPU thread walking the grid)."""
import math

import torch

from ..utils import MVNSqrt, mvn_loglikelihood, tria, whiten


def _sqrt_predict(F, cholQ, x: MVNSqrt):
    """filter.py:60-67"""
    return MVNSqrt(F @ x.mean, tria(torch.cat([F @ x.chol, cholQ], dim=1)))


def _sqrt_update(H, cholR, c, x: MVNSqrt):
    """filter.py:70-92 -> (posterior, log-likelihood increment, sigma^2 increment)"""
    m, cholP = x
    nx, ny = m.shape[0], c.shape[0]
    y_diff = -(H @ m + c)
    M = torch.cat([torch.cat([H @ cholP, cholR], dim=1),
                   torch.cat([cholP, torch.zeros((nx, ny), dtype=m.dtype, device=m.device)], dim=1)], dim=0)
    S = tria(M)
    I, G, cholP = S[:ny, :ny], S[ny:, :ny], S[ny:, ny:]
    wres = whiten(y_diff, I)
    ssq = torch.dot(wres, wres) / ny
    m = m + G @ torch.linalg.solve_triangular(I, y_diff.unsqueeze(-1), upper=False).squeeze(-1)
    return MVNSqrt(m, cholP), mvn_loglikelihood(y_diff, I), ssq


def _steps(dtm, n=None):
    F, QL = dtm.F, dtm.QL
    if F.dim() == 3:
        return [(F[k], QL[k]) for k in range(F.shape[0])]
    if n is None:
        raise ValueError("the transition model holds ONE (D,D) copy of F and QL (pof.convenience.set_up_solver does "
                         "not replicate them n times like the reference): pass the number of steps `n=`")
    return [(F, QL)] * int(n)


def _run_filter(x0, dtm, obs_at, n):
    x, ssq, ell = x0, 0.0, 0.0
    means, chols = [x0.mean], [x0.chol]
    steps = _steps(dtm, n)
    for k, (F, QL) in enumerate(steps):
        x = _sqrt_predict(F, QL, x)
        H, b, cholR = obs_at(k, x)
        x, ell_inc, ssq_inc = _sqrt_update(H, cholR, b, x)
        ssq, ell = ssq + ssq_inc, ell + ell_inc
        means.append(x.mean)
        chols.append(x.chol)
    return MVNSqrt(torch.stack(means), torch.stack(chols)), ell, ssq / len(steps)


def extended_kalman_filter(x0, discrete_transition_models, continuous_observation_model, *, n=None):
    """filter.py:9-30: EKF relinearised at every PREDICTED mean -> (filtered states (N), ell = +sum loglik, sigma^2)"""
    from ..observations import linearize

    return _run_filter(x0, discrete_transition_models, lambda k, x: linearize(continuous_observation_model, x), n)


def linear_noiseless_filter(x0, discrete_transition_models, discrete_observation_models):
    """filter.py:33-56: the same recursion for given affine observation models"""
    dom = discrete_observation_models
    n = dom.H.shape[0]
    zero = torch.zeros((dom.H.shape[1], dom.H.shape[1]), dtype=dom.H.dtype, device=dom.H.device)
    cholR = lambda k: zero if dom.cholR is None else dom.cholR[k]
    return _run_filter(x0, discrete_transition_models, lambda k, x: (dom.H[k], dom.b[k], cholR(k)), n)


def _sqrt_smooth(F, cholQ, xf: MVNSqrt, xs: MVNSqrt):
    """smoother.py:32-48"""
    nx = F.shape[0]
    Phi = tria(torch.cat([torch.cat([F @ xf.chol, cholQ], dim=1), torch.cat([xf.chol, torch.zeros_like(F)], dim=1)],
                         dim=0))
    Phi11, Phi21, Phi22 = Phi[:nx, :nx], Phi[nx:, :nx], Phi[nx:, nx:]
    gain = torch.linalg.solve_triangular(Phi11.T, Phi21.T, upper=True).T
    return MVNSqrt(xf.mean + gain @ (xs.mean - F @ xf.mean), tria(torch.cat([Phi22, gain @ xs.chol], dim=1)))


def smoothing(discrete_transition_models, filter_trajectory: MVNSqrt):
    """smoother.py:8-28 -> (smoothed states, obj); the objective keeps the reference's swapped arguments (quirk Q1)"""
    fm, fL = filter_trajectory
    n = fm.shape[0] - 1
    steps = _steps(discrete_transition_models, n)
    xs = MVNSqrt(fm[-1], fL[-1])
    means, chols = [xs.mean], [xs.chol]
    for k in range(n - 1, -1, -1):
        xs = _sqrt_smooth(steps[k][0], steps[k][1], MVNSqrt(fm[k], fL[k]), xs)
        means.append(xs.mean)
        chols.append(xs.chol)
    sm = torch.stack(means[::-1])
    obj = 0.0
    for k, (F, QL) in enumerate(steps):
        r = torch.linalg.solve_triangular(QL, (sm[k] - F @ sm[k + 1]).unsqueeze(-1), upper=False).squeeze(-1)
        obj = obj + torch.dot(r, r)
    return MVNSqrt(sm, torch.stack(chols[::-1])), obj
