"""CPU fp64 oracle for the parallel-in-time IEKS hot path of `pof`.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this module; only
`tests/`, `__graft_entry__.smoke()` and the CPU-baseline / reference legs of `bench.py`
may.  It is a NumPy/SymPy restatement of the reference's JAX formulas, function by
function, including the reference's quirks (swapped `obj` arguments, `whiten` solving
with L^T, double calibration, maxiters+1 iterations, LAPACK-signed unnormalised QR).

Parity status: the reference (JAX 0.3.23 + tornadox + diffrax) cannot be imported in
this environment, and its own tests hold no numerical golden vectors for this path.
The oracle is therefore PINNED against the reference's *published* results instead:
the `IEKS(3)_iterations` and `IEKS(3)_rmse_traj` columns of
`experiments/3_work_precision_diagram/data/*_Tesla_V100-SXM2-32GB.csv`
(see `tests/golden/published_ieks3.json`, `tests/test_oracle_published.py`).

Every function cites the reference file:line (relative to the reference root) it follows.
"""
from __future__ import annotations

import math
from typing import Callable, NamedTuple

import numpy as np
import scipy.linalg
import scipy.special


# --------------------------------------------------------------------------------------
# types  (pof/utils.py:9-11, pof/transitions.py:15-24, pof/observations.py:14-33)
# --------------------------------------------------------------------------------------
class MVNSqrt(NamedTuple):
    mean: np.ndarray
    chol: np.ndarray


class TransitionModel(NamedTuple):
    F: np.ndarray
    QL: np.ndarray


class AffineModel(NamedTuple):
    H: np.ndarray
    b: np.ndarray
    cholR: np.ndarray


class IVP(NamedTuple):
    name: str
    f: Callable  # f(t, y) -> (d,)   numpy
    jac: Callable  # jac(y) -> (d, d)
    y0: np.ndarray
    t0: float
    tmax: float
    taylor: Callable  # taylor(order) -> (order+1, d) derivatives y^{(k)}(t0)
    exprs: tuple = None  # the SymPy right-hand sides in the symbols y0..y{d-1} (oracle/threaded.py vectorises them)

    @property
    def t_span(self):
        return (self.t0, self.tmax)


# --------------------------------------------------------------------------------------
# small batched linear algebra  (pof/utils.py:22-41, 97-112)
# --------------------------------------------------------------------------------------
def tria(A):
    """pof/utils.py:33-41: tria(A) = qr(A^T, mode='r')^T, LAPACK Householder, signs kept."""
    R = np.linalg.qr(np.swapaxes(A, -1, -2), mode="r")
    return np.swapaxes(R, -1, -2)


def solve_lower(L, B, trans=False):
    """Batched solve_triangular(L, B, lower=True, trans=trans); L (...,n,n), B (...,n) or (...,n,k)."""
    vec = B.ndim == L.ndim - 1
    X = np.array(B[..., None] if vec else B, dtype=np.float64, copy=True)
    n = L.shape[-1]
    if not trans:
        for i in range(n):
            if i:
                X[..., i, :] -= np.einsum("...j,...jk->...k", L[..., i, :i], X[..., :i, :])
            X[..., i, :] /= L[..., i, i][..., None]
    else:  # solve L^T x = b (upper triangular system)
        for i in range(n - 1, -1, -1):
            if i < n - 1:
                X[..., i, :] -= np.einsum("...j,...jk->...k", L[..., i + 1 :, i], X[..., i + 1 :, :])
            X[..., i, :] /= L[..., i, i][..., None]
    return X[..., 0] if vec else X


def mvn_loglikelihood(x, chol_cov):
    """pof/utils.py:22-30 (batched)."""
    dim = chol_cov.shape[-1]
    y = solve_lower(chol_cov, x)
    diag = np.diagonal(chol_cov, axis1=-2, axis2=-1)
    normalizing_constant = np.sum(np.log(np.abs(diag)), -1) + dim * np.log(2 * np.pi) / 2.0
    norm_y = np.sum(y * y, -1)
    return -0.5 * norm_y - normalizing_constant


def whiten(m, cholP):
    """pof/utils.py:110-112: solve_triangular(cholP.T, m) with default lower=False,
    i.e. solves the UPPER system cholP^T x = m (quirk Q2)."""
    return solve_lower(cholP, m, trans=True)


def _mv(F, m):
    """rows of m times F^T; F is one (D,D) matrix or the reference's per-step stack (n,D,D) (vmapped models)."""
    return m @ F.T if F.ndim == 2 else np.einsum("nij,nj->ni", F, m)


def objective_function_value(mnext, m, F, QL):
    """pof/utils.py:97-101 (batched over leading axis; F, QL are single (D,D) or per-step (n,D,D))."""
    r = mnext - _mv(F, m)
    Lb = np.broadcast_to(QL, r.shape[:-1] + QL.shape[-2:])
    w = solve_lower(Lb, r)
    return np.sum(w * w, -1)


# --------------------------------------------------------------------------------------
# IWP prior  (pof/transitions.py:28-88)
# --------------------------------------------------------------------------------------
def preconditioned_discretize_1d(q):
    """pof/transitions.py:37-41."""
    A_1d = np.flip(scipy.linalg.pascal(q + 1, kind="lower", exact=False))
    Q_1d = np.flip(scipy.linalg.hilbert(q + 1))
    return A_1d, np.linalg.cholesky(Q_1d)


def preconditioned_discretize(d, q):
    """pof/transitions.py:44-50."""
    A_1d, L_Q1d = preconditioned_discretize_1d(q)
    Id = np.eye(d)
    return np.kron(Id, A_1d), np.kron(Id, L_Q1d)


def nordsieck_preconditioner(d, q, dt):
    """pof/transitions.py:53-68."""
    powers = np.arange(q, -1, -1)
    scales = scipy.special.factorial(powers)
    powers = powers + 0.5
    sv = (np.abs(dt) ** powers) / scales
    svi = (np.abs(dt) ** (-powers)) * scales
    Id = np.eye(d)
    return np.kron(Id, np.diag(sv)), np.kron(Id, np.diag(svi))


def projection_matrix(d, q, i):
    """pof/transitions.py:80-88."""
    return np.kron(np.eye(d), np.eye(1, q + 1, i))


# --------------------------------------------------------------------------------------
# set-up  (pof/convenience.py:13-45, 76-92; pof/initialization.py:15-22, 42-56)
# --------------------------------------------------------------------------------------
def taylor_mode_init(ivp: IVP, order):
    """pof/initialization.py:15-22: rows y^{(k)}(t0), k=0..q, flattened per dimension; chol=0.
    tornadox.init.TaylorMode is replaced by exact symbolic differentiation (IVP.taylor)."""
    derivs = ivp.taylor(order)  # (q+1, d)
    m0 = np.concatenate(list(derivs.T))  # [y1, y1', ..., y1^(q), y2, ...]
    D = m0.shape[0]
    return MVNSqrt(m0, np.zeros((D, D)))


def constant_init(ivp: IVP, order, N):
    """pof/initialization.py:42-56."""
    y0 = ivp.y0
    d = y0.shape[0]
    dy0 = ivp.f(None, y0)
    x0 = np.concatenate([y0[:, None], dy0[:, None], np.zeros((d, order - 1))], axis=1).reshape(1, -1)
    traj = np.repeat(x0, N, axis=0)
    D = traj.shape[1]
    return MVNSqrt(traj, np.zeros((N, D, D)))


def set_up_solver(ivp: IVP, ts, order):
    """pof/convenience.py:13-45.  F, QL are kept un-replicated (they are identical for every step)."""
    dt = (ts[1:] - ts[:-1])[0]
    d = ivp.y0.shape[0]
    F, QL = preconditioned_discretize(d, order)
    P, PI = nordsieck_preconditioner(d, order, dt)
    E0 = projection_matrix(d, order, 0) @ P
    E1 = projection_matrix(d, order, 1) @ P
    x0 = taylor_mode_init(ivp, order)
    x0 = MVNSqrt(PI @ x0.mean, PI @ x0.chol)
    return dict(ivp=ivp, ts=ts, dtm=TransitionModel(F, QL), x0=x0, E0=E0, E1=E1, P=P, PI=PI, order=order, d=d)


def coarse_ekf_init(ivp, order, ts, N=10):
    """pof/initialization.py:103-121: sequential EKS on a coarse grid of N points (full states, calibrated, mapped
    back with P), then piecewise-constant interpolation idx = floor(ts / coarse_dt) -- absolute times, i.e. t0 = 0 is
    assumed; JAX clamps out-of-range gather indices."""
    ts = np.asarray(ts, dtype=float)
    coarse_ts = np.linspace(ts[0], ts[-1], N)
    coarse_dt = coarse_ts[1] - coarse_ts[0]
    out, _ = sequential_eks_solve(ivp, coarse_ts, order, return_full_states=True)
    idxs = np.clip(np.floor(ts / coarse_dt).astype(int), 0, N - 1)
    return MVNSqrt(out.mean[idxs], out.chol[idxs])


def prior_init(ivp: IVP, order, ts):
    """pof/initialization.py:75-89 (`prior_init` -> `_prior_init`, :66-72): every row k >= 1 is ONE prediction of x0
    with the non-preconditioned model of step size ts[k] -- the ABSOLUTE time, not a grid spacing -- (quirk Q6:
    `discretize_transitions(iwp, steps=ts[1:])`, transitions.py:71-77, 91-99):
        mean_k = (P_k F PI_k) m0,   chol_k = tria([(P_k F PI_k) L0, P_k QL])   (sequential_filtsmooth/filter.py:60-67)
    and row 0 is x0 itself.  The result is in NON-preconditioned coordinates (convenience.py:88-90 applies no PI)."""
    ts = np.asarray(ts, dtype=np.float64)
    x0 = taylor_mode_init(ivp, order)
    d = ivp.y0.shape[0]
    F, QL = preconditioned_discretize(d, order)
    steps = ts[1:]
    n = steps.shape[0]
    D = x0.mean.shape[0]
    Fk = np.empty((n, D, D))
    QLk = np.empty((n, D, D))
    with np.errstate(divide="ignore", invalid="ignore"):
        for k in range(n):
            P, PI = nordsieck_preconditioner(d, order, steps[k])
            Fk[k] = P @ F @ PI
            QLk[k] = P @ QL
        means = np.einsum("nij,j->ni", Fk, x0.mean)
        finite = np.isfinite(QLk).all(axis=(1, 2)) & np.isfinite(Fk).all(axis=(1, 2))
        chols = np.full((n, D, D), np.nan)
        if finite.any():
            chols[finite] = tria(np.concatenate([Fk[finite] @ x0.chol, QLk[finite]], axis=-1))
    return MVNSqrt(np.concatenate([x0.mean[None], means]), np.concatenate([x0.chol[None], chols]))


def get_initial_trajectory(setup, method="constant"):
    """pof/convenience.py:76-92."""
    PI = setup["PI"]
    if method == "prior":
        return prior_init(setup["ivp"], setup["order"], setup["ts"])
    if method == "coarse":
        st = coarse_ekf_init(setup["ivp"], setup["order"], setup["ts"], N=100)
        return MVNSqrt(st.mean @ PI.T, np.einsum("ij,njk->nik", PI, st.chol))
    if method != "constant":
        raise NotImplementedError(method)
    N = len(setup["ts"])
    st = constant_init(setup["ivp"], setup["order"], N)
    return MVNSqrt(st.mean @ PI.T, np.einsum("ij,njk->nik", PI, st.chol))


# --------------------------------------------------------------------------------------
# linearisation  (pof/step.py:12-22, pof/observations.py:35-40)
# --------------------------------------------------------------------------------------
def linearize_at(setup, means):
    """Observation model x -> E1 x - f(E0 x) (convenience.py:28) linearised at each row of `means`:
    H = E1 - J_f(E0 m) E0 ; b = (E1 m - f(E0 m)) - H m ; cholR = 0."""
    ivp, E0, E1 = setup["ivp"], setup["E0"], setup["E1"]
    n = means.shape[0]
    d, D = E0.shape
    H = np.empty((n, d, D))
    b = np.empty((n, d))
    for k in range(n):
        m = means[k]
        y = E0 @ m
        Hk = E1 - ivp.jac(y) @ E0
        res = E1 @ m - ivp.f(None, y)
        H[k] = Hk
        b[k] = res - Hk @ m
    return AffineModel(H, b, np.zeros((n, d, d)))


# --------------------------------------------------------------------------------------
# jax.lax.associative_scan semantics [ext]  (call sites filter.py:35, smoother.py:15-17)
# --------------------------------------------------------------------------------------
def _interleave(even, odd):
    n = even.shape[0] + odd.shape[0]
    out = np.empty((n,) + even.shape[1:], dtype=even.dtype)
    out[0::2] = even
    out[1::2] = odd
    return out


def associative_scan(op, elems, reverse=False):
    """JAX's recursive odd/even schedule (the reference's association tree)."""
    if reverse:
        elems = tuple(np.flip(e, 0) for e in elems)
    n = elems[0].shape[0]

    def scan(e):
        m = e[0].shape[0]
        if m < 2:
            return e
        reduced = op(tuple(x[0:-1:2] for x in e), tuple(x[1::2] for x in e))
        odd = scan(reduced)
        if m % 2 == 0:
            even = op(tuple(x[:-1] for x in odd), tuple(x[2::2] for x in e))
        else:
            even = op(odd, tuple(x[2::2] for x in e))
        even = tuple(np.concatenate([x[:1], y]) for x, y in zip(e, even))
        return tuple(_interleave(a, b) for a, b in zip(even, odd))

    out = scan(tuple(elems))
    if reverse:
        out = tuple(np.flip(o, 0) for o in out)
    assert out[0].shape[0] == n
    return out


def sequential_scan(op, elems, reverse=False):
    """Left fold (an alternative, equally valid, association order) for cross-checks."""
    if reverse:
        elems = tuple(np.flip(e, 0) for e in elems)
    n = elems[0].shape[0]
    outs = [tuple(e[0:1] for e in elems)]
    for k in range(1, n):
        outs.append(op(outs[-1], tuple(e[k : k + 1] for e in elems)))
    out = tuple(np.concatenate([o[i] for o in outs]) for i in range(len(elems)))
    if reverse:
        out = tuple(np.flip(o, 0) for o in out)
    return out


# --------------------------------------------------------------------------------------
# parallel filter  (pof/parallel_filtsmooth/filter.py)
# --------------------------------------------------------------------------------------
def _T(x):
    return np.swapaxes(x, -1, -2)


def get_filter_elements(F, QL, H, c, cholR, ms, Ls):
    """filter.py:50-81 `_get_element`, batched over the leading axis."""
    n, ny, nx = H.shape
    m1 = _mv(F, ms)
    N1_ = tria(np.concatenate([F @ Ls, np.broadcast_to(QL, (n, nx, nx))], axis=-1))
    Psi_ = np.concatenate(
        [np.concatenate([H @ N1_, cholR], axis=-1), np.concatenate([N1_, np.zeros((n, nx, ny))], axis=-1)], axis=-2
    )
    Tria_Psi_ = tria(Psi_)
    Psi11 = Tria_Psi_[:, :ny, :ny]
    Psi21 = Tria_Psi_[:, ny:, :ny]
    U = Tria_Psi_[:, ny:, ny:]
    K = _T(solve_lower(Psi11, _T(Psi21), trans=True))
    HF = H @ F
    A = F - K @ HF
    b_sqr = m1 + np.einsum("nij,nj->ni", K, -np.einsum("nij,nj->ni", H, m1) - c)
    Z = _T(solve_lower(Psi11, HF))
    eta = np.einsum("nij,nj->ni", _T(solve_lower(Psi11, _T(Z), trans=True)), -c)
    if nx > ny:
        Z = np.concatenate([Z, np.zeros((n, nx, nx - ny))], axis=-1)
    else:
        Z = tria(Z)
    return A, b_sqr, U, eta, Z


def sqrt_filtering_operator(elem1, elem2):
    """filter.py:117-142, batched."""
    A1, b1, U1, eta1, Z1 = elem1
    A2, b2, U2, eta2, Z2 = elem2
    n, nx, _ = Z2.shape
    I = np.broadcast_to(np.eye(nx), (n, nx, nx))
    Xi = np.concatenate(
        [np.concatenate([_T(U1) @ Z2, I], axis=-1), np.concatenate([Z2, np.zeros_like(A1)], axis=-1)], axis=-2
    )
    tria_xi = tria(Xi)
    Xi11 = tria_xi[:, :nx, :nx]
    Xi21 = tria_xi[:, nx:, :nx]
    Xi22 = tria_xi[:, nx:, nx:]

    M = solve_lower(Xi11, _T(U1) @ _T(A2))
    A = A2 @ A1 - _T(M) @ _T(Xi21) @ A1
    m = solve_lower(Xi11, _T(U1))
    t = b1 + np.einsum("nij,nj->ni", U1 @ _T(U1), eta2)
    b = np.einsum("nij,nj->ni", A2 @ (I - _T(m) @ _T(Xi21)), t) + b2

    _U = _T(M)
    U = tria(np.concatenate([_U, U2], axis=-1))
    _e = solve_lower(Xi11, _T(Xi21), trans=True)
    t2 = eta2 - np.einsum("nij,nj->ni", Z2 @ _T(Z2), b1)
    eta = np.einsum("nij,nj->ni", _T(A1) @ (I - _T(_e) @ _T(U1)), t2) + eta1
    Z = tria(np.concatenate([_T(A1) @ Xi22, Z1], axis=-1))
    return A, b, U, eta, Z


def _get_obs(F, QL, H, c, cholR, m, cholP):
    """filter.py:84-93, batched."""
    n, ny, nx = H.shape
    predicted_mean = _mv(F, m)
    predicted_chol = tria(np.concatenate([F @ cholP, np.broadcast_to(QL, (n, nx, nx))], axis=-1))
    obs_mean = np.einsum("nij,nj->ni", H, predicted_mean) + c
    obs_chol = tria(np.concatenate([H @ predicted_chol, cholR], axis=-1))
    return obs_mean, obs_chol


def linear_noiseless_filtering(x0: MVNSqrt, dtm: TransitionModel, dom: AffineModel, scan=associative_scan):
    """filter.py:18-47."""
    F, QL = dtm
    H, c, cholR = dom
    n, d, D = H.shape
    ms = np.zeros((n, D))
    Ls = np.zeros((n, D, D))
    ms[0] = x0.mean
    Ls[0] = x0.chol
    elems = get_filter_elements(F, QL, H, c, cholR, ms, Ls)
    _, means, cholcovs, _, _ = scan(sqrt_filtering_operator, elems)
    means = np.concatenate([x0.mean[None], means])
    cholcovs = np.concatenate([x0.chol[None], cholcovs])

    obs_mean, obs_chol = _get_obs(F, QL, H, c, cholR, means[:-1], cholcovs[:-1])
    ress = whiten(obs_mean, obs_chol)  # filter.py:105-114 (quirks Q2, Q12)
    ssq = np.sum(ress * ress) / n / d
    nll = -np.sum(mvn_loglikelihood(obs_mean, obs_chol))  # filter.py:96-102
    obj = np.sum(objective_function_value(means[:-1], means[1:], F, QL))  # filter.py:43-45 (Q1)
    # sign-invariant variant, exported beside the reference's formula
    r2 = solve_lower(obs_chol, obs_mean)
    ssq_proper = np.sum(r2 * r2) / n / d
    return MVNSqrt(means, cholcovs), nll, obj, ssq, ssq_proper


# --------------------------------------------------------------------------------------
# parallel smoother  (pof/parallel_filtsmooth/smoother.py)
# --------------------------------------------------------------------------------------
def _sqrt_associative_params(F, QL, m, chol_P):
    """smoother.py:37-50, batched."""
    n, nx, _ = chol_P.shape
    QLb = np.broadcast_to(QL, (n, nx, nx))
    Phi = np.concatenate(
        [np.concatenate([F @ chol_P, QLb], axis=-1), np.concatenate([chol_P, np.zeros((n, nx, nx))], axis=-1)], axis=-2
    )
    Tria_Phi = tria(Phi)
    Phi11 = Tria_Phi[:, :nx, :nx]
    Phi21 = Tria_Phi[:, nx:, :nx]
    Dm = Tria_Phi[:, nx:, nx:]
    # smoother.py:48: E = solve(Phi11.T, Phi21.T).T  (general solve; Phi11 is triangular)
    E = _T(solve_lower(Phi11, _T(Phi21), trans=True))
    g = m - np.einsum("nij,nj->ni", E, _mv(F, m))
    return g, E, Dm


def sqrt_smoothing_operator(elem1, elem2):
    """smoother.py:53-63, batched."""
    g1, E1, D1 = elem1
    g2, E2, D2 = elem2
    g = np.einsum("nij,nj->ni", E2, g1) + g2
    E = E2 @ E1
    Dm = tria(np.concatenate([E2 @ D1, D2], axis=-1))
    return g, E, Dm


def smoothing(dtm: TransitionModel, filtered: MVNSqrt, scan=associative_scan):
    """smoother.py:8-34."""
    F, QL = dtm
    ms, Ps = filtered
    gs, Es, Ls = _sqrt_associative_params(F, QL, ms[:-1], Ps[:-1])
    gs = np.concatenate([gs, ms[-1][None]])
    Es = np.concatenate([Es, np.zeros_like(Ps[-1])[None]])
    Ls = np.concatenate([Ls, Ps[-1][None]])
    means, _, chols = scan(sqrt_smoothing_operator, (gs, Es, Ls), reverse=True)
    obj = np.sum(objective_function_value(means[:-1], means[1:], F, QL))  # smoother.py:20 (Q1)
    return MVNSqrt(means, chols), obj


def linear_filtsmooth(x0, dtm, dom, scan=associative_scan):
    """pof/parallel_filtsmooth/__init__.py:5-10.  Returns (states, nll, obj, ssq, ssq_proper)."""
    out, nll, _, ssq, ssq_proper = linear_noiseless_filtering(x0, dtm, dom, scan=scan)
    out, obj = smoothing(dtm, out, scan=scan)
    return out, nll, obj, ssq, ssq_proper


# --------------------------------------------------------------------------------------
# sequential filter / smoother  (pof/sequential_filtsmooth/*) -- baseline & cross-check
# --------------------------------------------------------------------------------------
def _sqrt_predict(F, QL, m, cholP):
    """sequential_filtsmooth/filter.py:60-67."""
    return F @ m, tria(np.concatenate([F @ cholP, QL], axis=1))


def _sqrt_update(H, cholR, c, m, cholP):
    """sequential_filtsmooth/filter.py:70-92."""
    nx, ny = m.shape[0], c.shape[0]
    y_diff = -(H @ m + c)
    M = np.block([[H @ cholP, cholR], [cholP, np.zeros((nx, ny))]])
    chol_S = tria(M)
    cholP_new = chol_S[ny:, ny:]
    G = chol_S[ny:, :ny]
    I_ = chol_S[:ny, :ny]
    wres = whiten(y_diff, I_)
    ssq = np.dot(wres, wres) / ny
    m_new = m + G @ solve_lower(I_, y_diff)
    ell = mvn_loglikelihood(y_diff, I_)
    return m_new, cholP_new, ell, ssq


def _sqrt_smooth(F, QL, mf, cholPf, ms, cholPs):
    """sequential_filtsmooth/smoother.py:32-48."""
    nx = F.shape[0]
    Phi = np.block([[F @ cholPf, QL], [cholPf, np.zeros_like(F)]])
    tria_Phi = tria(Phi)
    Phi11 = tria_Phi[:nx, :nx]
    Phi21 = tria_Phi[nx:, :nx]
    Phi22 = tria_Phi[nx:, nx:]
    gain = solve_lower(Phi11, Phi21.T, trans=True).T
    mean = mf + gain @ (ms - F @ mf)
    chol = tria(np.concatenate([Phi22, gain @ cholPs], axis=1))
    return mean, chol


def sequential_smoothing(dtm, filtered: MVNSqrt):
    """sequential_filtsmooth/smoother.py:8-28."""
    F, QL = dtm
    ms, Ps = filtered
    N = ms.shape[0]
    sm = np.empty_like(ms)
    sP = np.empty_like(Ps)
    sm[-1], sP[-1] = ms[-1], Ps[-1]
    for k in range(N - 2, -1, -1):
        sm[k], sP[k] = _sqrt_smooth(F, QL, ms[k], Ps[k], sm[k + 1], sP[k + 1])
    obj = np.sum(objective_function_value(sm[:-1], sm[1:], F, QL))
    return MVNSqrt(sm, sP), obj


def sequential_linear_filtsmooth(x0, dtm, dom):
    """sequential_filtsmooth/__init__.py:13-16 + filter.py:33-56 (returns ell = +sum loglik, quirk Q7)."""
    F, QL = dtm
    H, c, cholR = dom
    n = H.shape[0]
    D = x0.mean.shape[0]
    means = np.empty((n + 1, D))
    chols = np.empty((n + 1, D, D))
    means[0], chols[0] = x0
    m, P = x0
    ssq = 0.0
    ell = 0.0
    for k in range(n):
        m, P = _sqrt_predict(F, QL, m, P)
        m, P, e, s = _sqrt_update(H[k], cholR[k], c[k], m, P)
        ssq += s
        ell += e
        means[k + 1], chols[k + 1] = m, P
    ssq = ssq / n
    out, obj = sequential_smoothing(dtm, MVNSqrt(means, chols))
    return out, ell, obj, ssq


def sequential_eks(setup):
    """sequential_filtsmooth/__init__.py:5-10 + filter.py:9-30: EKF relinearised at the predicted mean."""
    F, QL = setup["dtm"]
    x0 = setup["x0"]
    n = len(setup["ts"]) - 1
    D = x0.mean.shape[0]
    means = np.empty((n + 1, D))
    chols = np.empty((n + 1, D, D))
    means[0], chols[0] = x0
    m, P = x0
    ssq = 0.0
    ell = 0.0
    for k in range(n):
        m, P = _sqrt_predict(F, QL, m, P)
        dom = linearize_at(setup, m[None])
        m, P, e, s = _sqrt_update(dom.H[0], dom.cholR[0], dom.b[0], m, P)
        ssq += s
        ell += e
        means[k + 1], chols[k + 1] = m, P
    ssq = ssq / n
    out, obj = sequential_smoothing(setup["dtm"], MVNSqrt(means, chols))
    return out, ell, obj, ssq


# --------------------------------------------------------------------------------------
# IEKS step, convergence, solve  (pof/step.py:33-45, pof/convergence_criteria.py, pof/solver.py)
# --------------------------------------------------------------------------------------
def ieks_step(setup, states: MVNSqrt, calibrate=True, sequential=False, scan=associative_scan):
    """step.py:33-45.  Returns (states, nll, obj, ssq, ssq_proper)."""
    dom = linearize_at(setup, states.mean[1:])
    if not sequential:
        out, nll, obj, ssq, ssqp = linear_filtsmooth(setup["x0"], setup["dtm"], dom, scan=scan)
    else:
        out, nll, obj, ssq = sequential_linear_filtsmooth(setup["x0"], setup["dtm"], dom)
        ssqp = float("nan")
    if calibrate:
        out = MVNSqrt(out.mean, math.sqrt(ssq) * out.chol if ssq >= 0 else np.nan * out.chol)
    return out, nll, obj, ssq, ssqp


def crit(obj, obj_old, nll, nll_old, means, means_old, rtol=1e-6, atol=1e-9):
    """convergence_criteria.py:4-13 (means isclose uses rtol=1e-13 and the default atol=1e-8)."""
    isnan = np.isnan(obj) or np.isnan(nll)
    obj_converged = np.isclose(obj_old, obj, rtol=rtol, atol=atol)
    means_converged = np.isclose(means_old, means, rtol=1e-13, atol=1e-8).all()
    return bool(isnan or obj_converged or means_converged)


def solve(ivp: IVP, ts, order, init="constant", calibrate=True, maxiters=10_000, sequential=False,
          scan=associative_scan, return_full=False):
    """solver.py:11-73."""
    setup = set_up_solver(ivp, ts, order)
    states = get_initial_trajectory(setup, method=init)
    nll = obj = ssq = 0.0
    nll_old = obj_old = 0.0
    states_old = states
    k = 0
    while True:
        first = k < 1
        if not first:
            converged = crit(obj, obj_old, nll, nll_old, states.mean, states_old.mean)
            if converged or not (k <= maxiters):
                break
        states_old, nll_old, obj_old = states, nll, obj
        states, nll, obj, ssq, _ = ieks_step(setup, states_old, sequential=sequential, scan=scan)
        k += 1
    info = dict(iterations=k, nll=nll, obj=obj, sigma_squared=ssq, calibrated=False)
    if calibrate:
        states = MVNSqrt(states.mean, math.sqrt(ssq) * states.chol)
        info["calibrated"] = True
    E0 = setup["E0"]
    ys = MVNSqrt(states.mean @ E0.T, np.einsum("ij,njk->nik", E0, states.chol))
    if return_full:
        return ys, info, states, setup
    return ys, info


def sequential_eks_solve(ivp: IVP, ts, order, return_full_states=False, calibrate=True):
    """solver.py:76-96."""
    setup = set_up_solver(ivp, ts, order)
    states, nll, obj, ssq = sequential_eks(setup)
    info = dict(nll=nll, obj=obj, sigma_squared=ssq, calibrated=False)
    if calibrate:
        states = MVNSqrt(states.mean, math.sqrt(ssq) * states.chol)
        info["calibrated"] = True
    M = setup["P"] if return_full_states else setup["E0"]
    return MVNSqrt(states.mean @ M.T, np.einsum("ij,njk->nik", M, states.chol)), info


# --------------------------------------------------------------------------------------
# regularised iteration  (pof/iterators.py:53-112)
# --------------------------------------------------------------------------------------
def qpm_ieks_iterator(setup, init_traj: MVNSqrt, reg_start=1e20, reg_final=1e-20, steps=40, tau_start=None,
                      tau_final=None, scan=associative_scan):
    """iterators.py:53-112 `_qpm_ieks_iterator`: quadratic-penalty IEKS.  Observation noise R = (reg / n)^2 I with
    cholR = reg / n * I (n = dtm.F.shape[0] transitions), reg and the tolerance tau shrink geometrically each time the
    stage has converged.  Yields (states, nll, obj, reg)."""
    x0, dtm = setup["x0"], setup["dtm"]
    dom = linearize_at(setup, init_traj.mean[1:])
    n, d = dom.b.shape
    reg_fact = (reg_final / reg_start) ** (1 / steps)
    if tau_start is None:
        tau_start, tau_final = 1e5, 1e-5
    tau_fact = (tau_final / tau_start) ** (1 / steps)
    reg, tau = reg_start, tau_start
    eye = np.broadcast_to(np.eye(d), (n, d, d))

    def fs(dom, reg):
        out, nll, obj, _, _ = linear_filtsmooth(x0, dtm, AffineModel(dom.H, dom.b, reg * eye / n), scan=scan)
        return out, nll, obj

    states, nll, obj = fs(dom, reg)
    yield states, nll, obj, reg
    while True:
        nll_old, obj_old, states_old = nll, obj, states
        dom = linearize_at(setup, states.mean[1:])
        states, nll, obj = fs(dom, reg)
        yield states, nll, obj, reg
        if crit(obj, obj_old, nll, nll_old, states.mean, states_old.mean, rtol=tau, atol=tau):
            reg *= reg_fact
            tau *= tau_fact
            if reg == 0:
                break
            elif reg < reg_final:
                reg = 0.0
                tau = min(tau_final, 1e-5)
        if np.isnan(nll) or np.isnan(obj):
            break


# --------------------------------------------------------------------------------------
# Levenberg-Marquardt-style iteration  (pof/iterators.py:109-133, pof/observations.py:65-83; upstream: uncalled, WIP)
# --------------------------------------------------------------------------------------
def stack_regularized(dom: AffineModel, means, reg):
    """observations.py:65-83 `linearize_regularized`, batched: [EK1 ; x ~ N(m, I / reg)], offsets `[f(m), -m]` as upstream
    (f(m) = b + H m for the EK1 model: b = f(m) - H m)."""
    n, d, D = dom.H.shape
    res = dom.b + np.einsum("nij,nj->ni", dom.H, means)
    eye = np.broadcast_to(np.eye(D), (n, D, D))
    cholR = np.zeros((n, d + D, d + D))
    cholR[:, d:, d:] = eye / np.sqrt(reg)
    return AffineModel(np.concatenate([dom.H, eye], axis=1), np.concatenate([res, -means], axis=1), cholR)


def lm_ieks_iterator(setup, init_traj: MVNSqrt, reg=1.0, scan=associative_scan):
    """iterators.py:109-133: yields (states, nll, obj, reg); `reg` stays constant (the accept / reject step is commented
    out upstream); stops when nll and obj are both isclose (rtol 1e-5, atol 1e-8) to the previous iterate's, or on NaN."""
    x0, dtm = setup["x0"], setup["dtm"]

    def one(states):
        means = states.mean[1:]
        out, nll, obj, _, _ = linear_filtsmooth(x0, dtm, stack_regularized(linearize_at(setup, means), means, reg),
                                                scan=scan)
        return out, nll, obj

    out, nll, obj = one(init_traj)
    yield out, nll, obj, reg
    while True:
        nll_old, obj_old = nll, obj
        out, nll, obj = one(out)
        yield out, nll, obj, reg
        if (np.isclose(nll_old, nll) and np.isclose(obj_old, obj)) or np.isnan(nll) or np.isnan(obj):
            break
