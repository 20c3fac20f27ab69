"""Sequential filter / smoother (reference pof/sequential_filtsmooth/): baseline and cross-check.

`linear_filtsmooth` is the same CUDA pass run with ONE chunk, i.e. a purely sequential square-root Kalman filter
and RTS smoother on a single thread (O(N) span) -- the sequential algorithm of the reference, not a parallel scan.
Returns `ell = +sum loglik` like the reference's sequential path (its sign differs from the parallel path's nll,
reference sequential_filtsmooth/filter.py:91 vs parallel_filtsmooth/filter.py:101).
"""
from ..parallel_filtsmooth import linear_filtsmooth as _pfs
from .loops import extended_kalman_filter, linear_noiseless_filter, smoothing  # noqa: F401  (reference __init__.py:1-2)


def linear_filtsmooth(x0, linear_transitions, linear_observations):
    n = linear_observations.H.shape[0]
    out, nll, obj, ssq = _pfs(x0, linear_transitions, linear_observations, chunk_len=n)
    return out, -nll, obj, ssq


def filtsmooth(x0, linear_transitions, continuous_observation_model, *, n=None):
    """reference sequential_filtsmooth/__init__.py:5-10: extended Kalman filter (relinearised at the predicted means)
    + RTS smoother -> (states, ell, obj, ssq).  For a built-in `pof.ivp` problem on the preconditioned model of
    `set_up_solver` (one (D,D) copy of F and QL: pass the number of steps `n`) this is the one-thread CUDA kernel
    `pof_sequential_eks_f64`; any other model runs the reference's step-by-step recursion in torch (`loops.py`)."""
    om, dtm = continuous_observation_model, linear_transitions
    lin = getattr(om.f, "_pof_lin", None)
    if lin is not None and lin["builtin"] is not None and dtm.F.dim() == 2 and n is not None:
        from ..transitions import IWP, preconditioned_discretize_1d
        from .eks import eks_filtsmooth

        _, qL = preconditioned_discretize_1d(IWP(num_derivatives=lin["q"], wiener_process_dimension=lin["d"]))
        setup = {"om": om, "dtm": dtm, "x0": x0, "ts": range(int(n) + 1), "_device": x0.mean.device, "_qL": qL}
        return eks_filtsmooth(setup)
    out, ell, ssq = extended_kalman_filter(x0, dtm, om, n=n)
    out, obj = smoothing(dtm, out)
    return out, ell, obj, ssq
