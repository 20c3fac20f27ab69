"""Extract the reference's published known answers for the IEKS path into a small JSON fixture.

Run ONCE in the build container (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_published.py
Source: experiments/3_work_precision_diagram/data/*_Tesla_V100-SXM2-32GB.csv, produced by the reference with
solve(init="constant", maxiters=1000), ts = linspace(t0, tmax, N)  (run_benchmark.py:63-80,114-125,179).
"""
import json
import os

import pandas as pd

REF = "/root/reference/experiments/3_work_precision_diagram/data"
FILES = {
    "logistic": "logistic_Tesla_V100-SXM2-32GB.csv",
    "fitzhughnagumo": "fhn_Tesla_V100-SXM2-32GB.csv",
    "vanderpol_mu1": "vdp0_Tesla_V100-SXM2-32GB.csv",
    "rigid_body": "rigidbody_Tesla_V100-SXM2-32GB.csv",
    "henonheiles_tmax10": "henonheiles_Tesla_V100-SXM2-32GB.csv",
}
out = {}
for name, fn in FILES.items():
    df = pd.read_csv(os.path.join(REF, fn))
    rows = []
    for _, r in df.iterrows():
        row = {"N": int(r["Ns"])}
        for q in (1, 2, 3):
            for col in ("iterations", "rmse_traj", "rmse_final", "runtime"):
                key = f"IEKS({q})_{col}"
                if key in df.columns and pd.notna(r[key]):
                    row[key] = float(r[key])
            for col in ("rmse_traj", "runtime"):
                key = f"EKS({q})_{col}"
                if key in df.columns and pd.notna(r[key]):
                    row[key] = float(r[key])
        rows.append(row)
    out[name] = rows
dst = os.path.join(os.path.dirname(__file__), "published_ieks3.json")
with open(dst, "w") as fh:
    json.dump(out, fh, indent=0)
print("wrote", dst, {k: len(v) for k, v in out.items()})
