"""`pof.iterators.lm_ieks_iterator` on the GPU against its oracle restatement, iterate by iterate.

The passes of this iterator (observation dimension d + D) run through torch's batched library calls
(`pof.parallel_filtsmooth.library_pass`, also checked against the oracle on CPU tensors in tests/test_library_pass.py);
the linearisation is the fused CUDA kernel."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ivps as oivps  # noqa: E402
from oracle import pof_oracle as O  # noqa: E402


@pytest.mark.parametrize("name,N,q", [("logistic", 30, 2), ("fitzhughnagumo", 64, 3)])
def test_lm_iterator_matches_oracle_iterates(native_lib, name, N, q):
    import pof.ivp
    from pof.convenience import get_initial_trajectory, set_up_solver
    from pof.iterators import lm_ieks_iterator

    ivp, oivp = getattr(pof.ivp, name)(), getattr(oivps, name)()
    tmax = 10.0
    ts = np.linspace(0.0, tmax, N)
    setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=ts, order=q)
    init = get_initial_trajectory(setup, method="constant")
    osetup = O.set_up_solver(oivp, ts, q)
    oinit = O.get_initial_trajectory(osetup)
    gen = lm_ieks_iterator(setup["dtm"], setup["om"], setup["x0"], init, reg=1.0)
    ogen = O.lm_ieks_iterator(osetup, oinit, reg=1.0)
    for k in range(4):
        st, nll, obj, reg = next(gen)
        ost, onll, oobj, oreg = next(ogen)
        assert reg == oreg == 1.0
        scale = np.abs(ost.mean).max(axis=0)
        assert (np.abs(st.mean.cpu().numpy() - ost.mean) <= 1e-7 * scale + 1e-10).all(), k
        assert abs(float(nll) - onll) <= 1e-7 * abs(onll) + 1e-9
        assert abs(float(obj) - oobj) <= 1e-7 * abs(oobj) + 1e-9
