"""Merge re-run columns of scripts/work_precision.py (e.g. --orders 4,5 after a fix) into the tracked CSVs.
    python scripts/merge_wp.py gpurun_out profiles"""
import csv, glob, os, sys
src, dst = sys.argv[1:3]
for f in sorted(glob.glob(os.path.join(src, "r01_work_precision_*.csv"))):
    g = os.path.join(dst, os.path.basename(f))
    new = {r["Ns"]: r for r in csv.DictReader(open(f))}
    if not os.path.exists(g):
        continue
    old = list(csv.DictReader(open(g)))
    cols = list(old[0].keys())
    for r in old:
        for k, v in new.get(r["Ns"], {}).items():
            if k not in cols:
                cols.append(k)
            r[k] = v
    cols = ["Ns"] + sorted(c for c in cols if c != "Ns")
    with open(g, "w", newline="") as fh:
        w = csv.DictWriter(fh, fieldnames=cols)
        w.writeheader()
        for r in old:
            w.writerow(r)
    print("merged", f, "->", g)
