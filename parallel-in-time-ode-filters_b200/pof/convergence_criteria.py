"""IEKS stopping rule (reference pof/convergence_criteria.py:4-13).

The elementwise `isclose(means_old, means, rtol=1e-13)` reduction over the full (N, D) state is computed inside the
smoother kernel (scalar POF_S_NOT_CLOSE); this module holds the scalar logic.
"""
import math


def crit_scalars(obj, obj_old, nll, nll_old, n_not_close, *, rtol=1e-6, atol=1e-9):
    isnan = math.isnan(obj) or math.isnan(nll)
    # numpy/jax isclose(a=obj_old, b=obj): |a-b| <= atol + rtol*|b|
    obj_converged = abs(obj_old - obj) <= atol + rtol * abs(obj) if not isnan else False
    means_converged = n_not_close == 0
    return bool(isnan or obj_converged or means_converged)


def crit(obj, obj_old, nll, nll_old, states, states_old, *, rtol=1e-6, atol=1e-9):
    """Reference signature; `states` hold torch tensors."""
    import torch

    m, mo = states.mean, states_old.mean
    close = (m - mo).abs() <= 1e-8 + 1e-13 * m.abs()
    return crit_scalars(float(obj), float(obj_old), float(nll), float(nll_old), int((~close).sum().item()),
                        rtol=rtol, atol=atol)
