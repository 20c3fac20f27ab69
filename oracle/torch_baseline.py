"""GPU *library* baseline: the reference's parallel IEKS pass restated with stock torch-CUDA library calls.

TEST / BENCH INFRASTRUCTURE ONLY (see pof_oracle.py): not imported by the product package.

The reference runs on a GPU through XLA's stock lowering of `jnp.linalg.qr`, `solve_triangular`, small matmuls and
`jax.lax.associative_scan`.  JAX is not installable in this image, so the closest stand-in for "the
reference's single-GPU JAX time" on the same B200 is the same algorithm issued through the same KIND of library calls:
batched `torch.linalg.qr` (cuSOLVER / MAGMA), `torch.linalg.solve_triangular` (cuBLAS trsm), batched matmuls, and the
recursive odd/even scan of `jax.lax.associative_scan`, every level a handful of library launches with all
intermediates through HBM.  Function by function it follows oracle/pof_oracle.py, i.e. the reference's
  pof/parallel_filtsmooth/filter.py:18-142, smoother.py:8-63, pof/utils.py:22-41,97-112.
None of this repo's kernels are involved.  Validated against the NumPy oracle on the CPU (tests/test_torch_baseline.py).
"""
from __future__ import annotations

import math

import torch


def _T(x):
    return x.transpose(-1, -2)


def tria(A):
    """pof/utils.py:33-41"""
    return _T(torch.linalg.qr(_T(A), mode="r").R)


def solve_lower(L, B, trans=False):
    """solve_triangular(L, B, lower=True, trans=trans), batched; B (..., n) or (..., n, k)"""
    vec = B.dim() == L.dim() - 1
    X = B.unsqueeze(-1) if vec else B
    if trans:
        out = torch.linalg.solve_triangular(_T(L), X, upper=True)
    else:
        out = torch.linalg.solve_triangular(L, X, upper=False)
    return out.squeeze(-1) if vec else out


def _mv(F, m):
    return m @ F.T if F.dim() == 2 else torch.einsum("nij,nj->ni", F, m)


def _interleave(even, odd):
    n = even.shape[0] + odd.shape[0]
    out = torch.empty((n,) + tuple(even.shape[1:]), dtype=even.dtype, device=even.device)
    out[0::2] = even
    out[1::2] = odd
    return out


def associative_scan(op, elems, reverse=False):
    """jax.lax.associative_scan (recursive odd/even), as in oracle.pof_oracle.associative_scan"""
    if reverse:
        elems = tuple(torch.flip(e, dims=(0,)) for e in elems)
        out = associative_scan(op, elems)
        return tuple(torch.flip(e, dims=(0,)) for e in out)
    n = elems[0].shape[0]
    if n < 2:
        return elems
    reduced = op(tuple(e[0:n - 1:2] for e in elems), tuple(e[1::2] for e in elems))
    odd = associative_scan(op, reduced)
    if n % 2 == 0:
        even = op(tuple(e[:-1] for e in odd), tuple(e[2::2] for e in elems))
    else:
        even = op(odd, tuple(e[2::2] for e in elems))
    even = tuple(torch.cat([e[0:1], ev]) for e, ev in zip(elems, even))
    return tuple(_interleave(ev, od) for ev, od in zip(even, odd))


def get_filter_elements(F, QL, H, c, cholR, ms, Ls):
    """filter.py:50-81"""
    n, ny, nx = H.shape
    m1 = _mv(F, ms)
    N1_ = tria(torch.cat([F @ Ls, QL.expand(n, nx, nx)], dim=-1))
    Psi_ = torch.cat([torch.cat([H @ N1_, cholR], dim=-1),
                      torch.cat([N1_, torch.zeros((n, nx, ny), dtype=H.dtype, device=H.device)], dim=-1)], dim=-2)
    Tria_Psi_ = tria(Psi_)
    Psi11 = Tria_Psi_[:, :ny, :ny]
    Psi21 = Tria_Psi_[:, ny:, :ny]
    U = Tria_Psi_[:, ny:, ny:]
    K = _T(solve_lower(Psi11, _T(Psi21), trans=True))
    HF = H @ F
    A = F - K @ HF
    b_sqr = m1 + torch.einsum("nij,nj->ni", K, -torch.einsum("nij,nj->ni", H, m1) - c)
    Z = _T(solve_lower(Psi11, HF))
    eta = torch.einsum("nij,nj->ni", _T(solve_lower(Psi11, _T(Z), trans=True)), -c)
    if nx > ny:
        Z = torch.cat([Z, torch.zeros((n, nx, nx - ny), dtype=H.dtype, device=H.device)], dim=-1)
    else:
        Z = tria(Z)
    return A, b_sqr, U.contiguous(), eta, Z


def sqrt_filtering_operator(elem1, elem2):
    """filter.py:117-142"""
    A1, b1, U1, eta1, Z1 = elem1
    A2, b2, U2, eta2, Z2 = elem2
    n, nx, _ = Z2.shape
    I = torch.eye(nx, dtype=A1.dtype, device=A1.device).expand(n, nx, nx)
    Xi = torch.cat([torch.cat([_T(U1) @ Z2, I], dim=-1), torch.cat([Z2, torch.zeros_like(A1)], dim=-1)], dim=-2)
    tria_xi = tria(Xi)
    Xi11 = tria_xi[:, :nx, :nx]
    Xi21 = tria_xi[:, nx:, :nx]
    Xi22 = tria_xi[:, nx:, nx:]
    M = solve_lower(Xi11, _T(U1) @ _T(A2))
    A = A2 @ A1 - _T(M) @ _T(Xi21) @ A1
    m = solve_lower(Xi11, _T(U1))
    t = b1 + torch.einsum("nij,nj->ni", U1 @ _T(U1), eta2)
    b = torch.einsum("nij,nj->ni", A2 @ (I - _T(m) @ _T(Xi21)), t) + b2
    U = tria(torch.cat([_T(M), U2], dim=-1))
    _e = solve_lower(Xi11, _T(Xi21), trans=True)
    t2 = eta2 - torch.einsum("nij,nj->ni", Z2 @ _T(Z2), b1)
    eta = torch.einsum("nij,nj->ni", _T(A1) @ (I - _T(_e) @ _T(U1)), t2) + eta1
    Z = tria(torch.cat([_T(A1) @ Xi22, Z1], dim=-1))
    return A, b, U, eta, Z


def _get_obs(F, QL, H, c, cholR, m, cholP):
    """filter.py:84-93"""
    n, ny, nx = H.shape
    predicted_mean = _mv(F, m)
    predicted_chol = tria(torch.cat([F @ cholP, QL.expand(n, nx, nx)], dim=-1))
    obs_mean = torch.einsum("nij,nj->ni", H, predicted_mean) + c
    obs_chol = tria(torch.cat([H @ predicted_chol, cholR], dim=-1))
    return obs_mean, obs_chol


def objective_function_value(mnext, m, F, QL):
    """utils.py:97-101"""
    r = mnext - _mv(F, m)
    w = solve_lower(QL.expand(r.shape[0], *QL.shape[-2:]), r)
    return (w * w).sum(-1)


def linear_noiseless_filtering(x0m, x0c, F, QL, H, c, cholR):
    """filter.py:18-47 -> (means, chols, nll, ssq, ssq_proper)"""
    n, d, D = H.shape
    ms = torch.zeros((n, D), dtype=H.dtype, device=H.device)
    Ls = torch.zeros((n, D, D), dtype=H.dtype, device=H.device)
    ms[0] = x0m
    Ls[0] = x0c
    elems = get_filter_elements(F, QL, H, c, cholR, ms, Ls)
    _, means, cholcovs, _, _ = associative_scan(sqrt_filtering_operator, elems)
    means = torch.cat([x0m[None], means])
    cholcovs = torch.cat([x0c[None], cholcovs])
    obs_mean, obs_chol = _get_obs(F, QL, H, c, cholR, means[:-1], cholcovs[:-1])
    ress = solve_lower(obs_chol, obs_mean, trans=True)  # whiten: utils.py:110-112 (solves with L^T)
    ssq = (ress * ress).sum() / n / d
    y = solve_lower(obs_chol, obs_mean)
    diag = torch.diagonal(obs_chol, dim1=-2, dim2=-1)
    ll = -0.5 * (y * y).sum(-1) - (diag.abs().log().sum(-1) + d * math.log(2 * math.pi) / 2.0)
    nll = -ll.sum()
    ssq_proper = (y * y).sum() / n / d
    return means, cholcovs, nll, ssq, ssq_proper


def _sqrt_associative_params(F, QL, m, chol_P):
    """smoother.py:37-50"""
    n, nx, _ = chol_P.shape
    Phi = torch.cat([torch.cat([F @ chol_P, QL.expand(n, nx, nx)], dim=-1),
                     torch.cat([chol_P, torch.zeros((n, nx, nx), dtype=m.dtype, device=m.device)], dim=-1)], dim=-2)
    Tria_Phi = tria(Phi)
    Phi11 = Tria_Phi[:, :nx, :nx]
    Phi21 = Tria_Phi[:, nx:, :nx]
    Dm = Tria_Phi[:, nx:, nx:]
    # smoother.py:48 is a GENERAL solve (LU) in the reference: jlinalg.solve(Phi11.T, Phi21.T).T
    E = _T(torch.linalg.solve(_T(Phi11), _T(Phi21)))
    g = m - torch.einsum("nij,nj->ni", E, _mv(F, m))
    return g, E, Dm.contiguous()


def sqrt_smoothing_operator(elem1, elem2):
    """smoother.py:53-63"""
    g1, E1, D1 = elem1
    g2, E2, D2 = elem2
    g = torch.einsum("nij,nj->ni", E2, g1) + g2
    E = E2 @ E1
    Dm = tria(torch.cat([E2 @ D1, D2], dim=-1))
    return g, E, Dm


def smoothing(F, QL, ms, Ps):
    """smoother.py:8-34"""
    gs, Es, Ls = _sqrt_associative_params(F, QL, ms[:-1], Ps[:-1])
    gs = torch.cat([gs, ms[-1][None]])
    Es = torch.cat([Es, torch.zeros_like(Ps[-1])[None]])
    Ls = torch.cat([Ls, Ps[-1][None]])
    means, _, chols = associative_scan(sqrt_smoothing_operator, (gs, Es, Ls), reverse=True)
    obj = objective_function_value(means[:-1], means[1:], F, QL).sum()
    return means, chols, obj


def linearize(f_and_jac, E0, E1, means):
    """step.py:12-22 / observations.py:35-40 with a batched torch vector field: f_and_jac(Y (n,d)) -> (f (n,d), J (n,d,d))"""
    Y = means @ E0.T
    fv, J = f_and_jac(Y)
    H = E1[None] - J @ E0[None]
    res = means @ E1.T - fv
    b = res - torch.einsum("nij,nj->ni", H, means)
    n, d = b.shape
    return H, b, torch.zeros((n, d, d), dtype=means.dtype, device=means.device)


def ieks_step(f_and_jac, E0, E1, F, QL, x0m, x0c, means, calibrate=True):
    """step.py:33-45 -> (means, chols, nll, obj, ssq, ssq_proper)"""
    H, c, cholR = linearize(f_and_jac, E0, E1, means[1:])
    fm, fc, nll, ssq, ssqp = linear_noiseless_filtering(x0m, x0c, F, QL, H, c, cholR)
    sm, sc, obj = smoothing(F, QL, fm, fc)
    if calibrate:
        sc = torch.sqrt(ssq) * sc
    return sm, sc, nll, obj, ssq, ssqp


def fhn_f_and_jac(p=(0.7, 0.8, 1 / 12.5, 0.5)):
    """FitzHugh-Nagumo (pof/ivp.py:44-60), batched"""
    a, b, tinv, l = p

    def fj(Y):
        v, w = Y[:, 0], Y[:, 1]
        f = torch.stack([v - v ** 3 / 3 - w + l, tinv * (v + a - b * w)], dim=1)
        J = torch.zeros((Y.shape[0], 2, 2), dtype=Y.dtype, device=Y.device)
        J[:, 0, 0] = 1 - v ** 2
        J[:, 0, 1] = -1.0
        J[:, 1, 0] = tinv
        J[:, 1, 1] = -tinv * b
        return f, J

    return fj
