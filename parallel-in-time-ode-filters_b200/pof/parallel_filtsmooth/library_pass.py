"""Filter + smoother pass for an ARBITRARY observation dimension through torch's batched library calls.

The CUDA kernels of this package take observations of the ODE dimension d (dy = d <= D).  The reference's seam
`linear_filtsmooth(x0, dtm, dom)` (parallel_filtsmooth/__init__.py:5-10) has no such restriction, and one upstream
caller needs more: `lm_ieks_iterator` (iterators.py:109-133) stacks the EK1 model with a pseudo-observation of the
whole state (`linearize_regularized`, observations.py:65-83), dy = d + D > D.  For that case -- and only for it -- the
pass runs here as batched `torch.linalg.qr` / `solve_triangular` / matmul calls on the tensors' device with the
reference's own scan (`jax.lax.associative_scan`'s recursive odd/even schedule), i.e. what XLA does for the reference
on a GPU.  It is a functional path, not a fast one (~150x the kernels' time per pass at N = 2^20);
nothing in `solve` or the hot path uses it.

Formulas: parallel_filtsmooth/filter.py:18-142, smoother.py:8-63, utils.py:22-41,97-112 of the reference.
"""
import math

import torch

from ..utils import MVNSqrt


def _T(x):
    return x.transpose(-1, -2)


def _tria(A):
    """utils.py:33-41: lower-triangular L with L L^T = A A^T"""
    return _T(torch.linalg.qr(_T(A), mode="r").R)


def _solve_lower(L, B, trans=False):
    """L x = B (or L^T x = B), batched; B is a batch of vectors or of matrices"""
    vec = B.dim() == L.dim() - 1
    X = B.unsqueeze(-1) if vec else B
    X = torch.linalg.solve_triangular(_T(L), X, upper=True) if trans else torch.linalg.solve_triangular(L, X, upper=False)
    return X.squeeze(-1) if vec else X


def _mv(M, v):
    return torch.einsum("...ij,...j->...i", M, v)


def _scan(op, elems, reverse=False):
    """jax.lax.associative_scan: reduce neighbouring pairs, recurse, fill in the even positions"""
    if reverse:
        out = _scan(op, tuple(torch.flip(e, dims=(0,)) for e in elems))
        return tuple(torch.flip(e, dims=(0,)) for e in out)
    n = elems[0].shape[0]
    if n < 2:
        return elems
    odd = _scan(op, op(tuple(e[0:n - 1:2] for e in elems), tuple(e[1::2] for e in elems)))
    if n % 2 == 0:
        even = op(tuple(e[:-1] for e in odd), tuple(e[2::2] for e in elems))
    else:
        even = op(odd, tuple(e[2::2] for e in elems))
    outs = []
    for e, ev, od in zip(elems, even, odd):
        full = torch.empty_like(e)
        full[0] = e[0]
        full[2::2] = ev
        full[1::2] = od
        outs.append(full)
    return tuple(outs)


def _filter_elements(F, QL, H, c, cholR, m0, L0):
    """filter.py:50-81, all steps at once; step 0 carries the initial state, the others a zero prior"""
    n, ny, nx = H.shape
    ms = torch.zeros((n, nx), dtype=H.dtype, device=H.device)
    Ls = torch.zeros((n, nx, nx), dtype=H.dtype, device=H.device)
    ms[0], Ls[0] = m0, L0
    m1 = _mv(F, ms)
    N1 = _tria(torch.cat([F @ Ls, QL], dim=-1))
    zeros = torch.zeros((n, nx, ny), dtype=H.dtype, device=H.device)
    Psi = _tria(torch.cat([torch.cat([H @ N1, cholR], dim=-1), torch.cat([N1, zeros], dim=-1)], dim=-2))
    Psi11, Psi21, U = Psi[:, :ny, :ny], Psi[:, ny:, :ny], Psi[:, ny:, ny:]
    K = _T(_solve_lower(Psi11, _T(Psi21), trans=True))
    HF = H @ F
    A = F - K @ HF
    b = m1 + _mv(K, -_mv(H, m1) - c)
    Z = _T(_solve_lower(Psi11, HF))
    eta = _mv(_T(_solve_lower(Psi11, _T(Z), trans=True)), -c)
    if nx > ny:
        Z = torch.cat([Z, torch.zeros((n, nx, nx - ny), dtype=H.dtype, device=H.device)], dim=-1)
    else:
        Z = _tria(Z)
    return A, b, U.contiguous(), eta, Z


def _filter_op(e1, e2):
    """filter.py:117-142 (e1 earlier in time)"""
    A1, b1, U1, eta1, Z1 = e1
    A2, b2, U2, eta2, Z2 = e2
    n, nx, _ = Z2.shape
    eye = torch.eye(nx, dtype=A1.dtype, device=A1.device).expand(n, nx, nx)
    Xi = _tria(torch.cat([torch.cat([_T(U1) @ Z2, eye], dim=-1), torch.cat([Z2, torch.zeros_like(A1)], dim=-1)], dim=-2))
    Xi11, Xi21, Xi22 = Xi[:, :nx, :nx], Xi[:, nx:, :nx], Xi[:, nx:, nx:]
    M = _solve_lower(Xi11, _T(U1) @ _T(A2))
    A = A2 @ A1 - _T(M) @ _T(Xi21) @ A1
    mm = _solve_lower(Xi11, _T(U1))
    b = _mv(A2 @ (eye - _T(mm) @ _T(Xi21)), b1 + _mv(U1 @ _T(U1), eta2)) + b2
    U = _tria(torch.cat([_T(M), U2], dim=-1))
    ee = _solve_lower(Xi11, _T(Xi21), trans=True)
    eta = _mv(_T(A1) @ (eye - _T(ee) @ _T(U1)), eta2 - _mv(Z2 @ _T(Z2), b1)) + eta1
    Z = _tria(torch.cat([_T(A1) @ Xi22, Z1], dim=-1))
    return A, b, U, eta, Z


def _smooth_op(e1, e2):
    """smoother.py:53-63 (with reverse=True, e1 is the later, accumulated element)"""
    g1, E1, D1 = e1
    g2, E2, D2 = e2
    return _mv(E2, g1) + g2, E2 @ E1, _tria(torch.cat([E2 @ D1, D2], dim=-1))


def _rep(M, n):
    return M if M.dim() == 3 else M.unsqueeze(0).expand(n, *M.shape)


def linear_noiseless_filtering_library(x0: MVNSqrt, dtm, dom):
    """`linear_noiseless_filtering(x0, dtm, dom)` (filter.py:18-47) for any state / observation dimension
    -> (filtered MVNSqrt (N,D), (N,D,D), nll, obj, ssq)"""
    H, c, cholR = dom.H, dom.b, dom.cholR
    n, ny, nx = H.shape
    F, QL = _rep(dtm.F, n), _rep(dtm.QL, n)
    if cholR is None:
        cholR = torch.zeros((n, ny, ny), dtype=H.dtype, device=H.device)
    _, fm, fL, _, _ = _scan(_filter_op, _filter_elements(F, QL, H, c, cholR, x0.mean, x0.chol))
    fm = torch.cat([x0.mean[None], fm])
    fL = torch.cat([x0.chol[None], fL])
    # innovation statistics from the filtered states (filter.py:84-114; `whiten` solves with L^T, utils.py:110-112)
    pm = _mv(F, fm[:-1])
    pL = _tria(torch.cat([F @ fL[:-1], QL], dim=-1))
    om = _mv(H, pm) + c
    oL = _tria(torch.cat([H @ pL, cholR], dim=-1))
    w = _solve_lower(oL, om, trans=True)
    ssq = (w * w).sum() / n / ny
    y = _solve_lower(oL, om)
    logdet = torch.diagonal(oL, dim1=-2, dim2=-1).abs().log().sum(-1)
    nll = (0.5 * (y * y).sum(-1) + logdet + ny * math.log(2 * math.pi) / 2.0).sum()
    # the filter's objective with the reference's swapped arguments (filter.py:43-45, utils.py:97-101)
    wq = _solve_lower(QL, fm[:-1] - _mv(F, fm[1:]))
    return MVNSqrt(fm, fL), nll, (wq * wq).sum(), ssq


def smoothing_library(dtm, filtered: MVNSqrt):
    """`smoothing(dtm, filtered)` (smoother.py:8-50): elements from the joint triangularisation, suffix scan
    -> (smoothed MVNSqrt, obj)"""
    fm, fL = filtered.mean, filtered.chol
    n, nx = fm.shape[0] - 1, fm.shape[1]
    F, QL = _rep(dtm.F, n), _rep(dtm.QL, n)
    zeros = torch.zeros((n, nx, nx), dtype=fm.dtype, device=fm.device)
    Phi = _tria(torch.cat([torch.cat([F @ fL[:-1], QL], dim=-1), torch.cat([fL[:-1], zeros], dim=-1)], dim=-2))
    Phi11, Phi21, Dm = Phi[:, :nx, :nx], Phi[:, nx:, :nx], Phi[:, nx:, nx:]
    E = _T(torch.linalg.solve(_T(Phi11), _T(Phi21)))  # smoother.py:48 is a general solve upstream
    g = fm[:-1] - _mv(E, _mv(F, fm[:-1]))
    gs = torch.cat([g, fm[-1:]])
    Es = torch.cat([E, torch.zeros_like(fL[-1:])])
    Ds = torch.cat([Dm, fL[-1:]])
    sm, _, sL = _scan(_smooth_op, (gs, Es, Ds), reverse=True)
    # objective with the reference's swapped arguments (smoother.py:20, utils.py:97-101)
    wq = _solve_lower(QL, sm[:-1] - _mv(F, sm[1:]))
    return MVNSqrt(sm, sL), (wq * wq).sum()


def linear_filtsmooth_library(x0: MVNSqrt, dtm, dom):
    """`linear_filtsmooth(x0, dtm, dom)` (parallel_filtsmooth/__init__.py:5-10) for any observation dimension:
    -> (MVNSqrt(means (N,D), chols (N,D,D)), nll, obj, ssq), uncalibrated, like the reference."""
    filtered, nll, _, ssq = linear_noiseless_filtering_library(x0, dtm, dom)
    out, obj = smoothing_library(dtm, filtered)
    return out, nll, obj, ssq
