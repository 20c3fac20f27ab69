"""The GPU-library baseline arm of bench.py (oracle/torch_baseline.py: the reference's pass through stock torch library
calls) is the same algorithm as the NumPy oracle: checked here on the CPU device at a size that takes seconds."""
import numpy as np
import torch

from oracle import ivps as oivps
from oracle import pof_oracle as O
from oracle import torch_baseline as TB


def test_torch_baseline_matches_numpy_oracle():
    ivp = oivps.fitzhughnagumo()
    N = 333
    ts = np.linspace(0, 100, N)
    s = O.set_up_solver(ivp, ts, 3)
    st = O.get_initial_trajectory(s)
    out, nll, obj, ssq, ssqp = O.ieks_step(s, st, calibrate=False)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64)
    sm, sc, tnll, tobj, tssq, tssqp = TB.ieks_step(TB.fhn_f_and_jac(), t(s["E0"]), t(s["E1"]), t(s["dtm"].F),
                                                   t(s["dtm"].QL), t(s["x0"].mean), t(s["x0"].chol), t(st.mean),
                                                   calibrate=False)
    E0 = s["E0"]
    y, yo = sm.numpy() @ E0.T, out.mean @ E0.T
    assert np.abs(y - yo).max() <= 1e-10 * np.abs(yo).max()
    C = (sc @ sc.transpose(-1, -2)).numpy()
    Co = out.chol @ np.swapaxes(out.chol, -1, -2)
    assert np.abs(C - Co).max() <= 1e-9 * np.abs(Co).max()
    assert abs(float(tnll) - nll) <= 1e-10 * abs(nll) and abs(float(tobj) - obj) <= 1e-10 * abs(obj)
    assert abs(float(tssqp) - ssqp) <= 1e-9 * ssqp and abs(float(tssq) - ssq) <= 1e-2 * ssq
