"""Sequential filter / smoother (reference pof/sequential_filtsmooth/): baseline and cross-check.

`linear_filtsmooth` is the same CUDA pass run with ONE chunk, i.e. a purely sequential square-root Kalman filter
and RTS smoother on a single thread (O(N) span) -- the sequential algorithm of the reference, not a parallel scan.
Returns `ell = +sum loglik` like the reference's sequential path (its sign differs from the parallel path's nll,
reference sequential_filtsmooth/filter.py:91 vs parallel_filtsmooth/filter.py:101).
"""
from ..parallel_filtsmooth import linear_filtsmooth as _pfs


def linear_filtsmooth(x0, linear_transitions, linear_observations):
    n = linear_observations.H.shape[0]
    out, nll, obj, ssq = _pfs(x0, linear_transitions, linear_observations, chunk_len=n)
    return out, -nll, obj, ssq
