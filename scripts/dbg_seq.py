import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "parallel-in-time-ode-filters_b200")]
import numpy as np, torch
import pof.ivp
from pof.convenience import set_up_solver
from pof.sequential_filtsmooth.eks import eks_filtsmooth
from oracle import ivps as oivps, pof_oracle as O
ivp = pof.ivp.logistic(); oivp = oivps.logistic()
ts = np.linspace(0, 10, 21)
setup = set_up_solver(f=ivp.f, y0=ivp.y0, ts=ts, order=3)
st, ell, obj, ssq = eks_filtsmooth(setup)
torch.cuda.synchronize()
osetup = O.set_up_solver(oivp, ts, 3)
ost, oell, oobj, ossq = O.sequential_eks(osetup)
print("ell", ell, oell, "obj", obj, oobj, "ssq", ssq, ossq)
print(np.abs(st.mean.cpu().numpy() - ost.mean).max(axis=1))
