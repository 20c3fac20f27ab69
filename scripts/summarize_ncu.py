"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_ncu.py <launch_list.csv> <full.ncu-rep> <tag> [<more.ncu-rep> ...]

Also writes profiles/r02_ncu_kernels.json (dram bytes, FP64-pipe / LSU / issue utilisation per kernel) together with
the hash of the kernel sources the capture was taken from: bench.py reports these counters (`roofline.frac`,
`roofline.traffic`) only while that hash equals the hash of the sources the running library was built from.
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

launches, rep, tag = sys.argv[1:4]
more_reps = sys.argv[4:]
out_dir = "profiles"

# ---- launch list: per-kernel share of one step
lines = [l for l in open(launches) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
names = [r["Kernel Name"] for r in rows]
folds = [i for i, n in enumerate(names) if "fold" in n]
start, end = (folds[-2] - 1, folds[-1] - 1) if len(folds) >= 2 else (0, len(rows))
agg = collections.OrderedDict()
for r in rows[start:end]:
    n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
with open(f"{out_dir}/{tag}_launch_shares.md", "w") as fh:
    fh.write(f"# {tag}: kernels of ONE IEKS iteration (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n")
    fh.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
    fh.write("| kernel | launches | total us | share |\n|---|---|---|---|\n")
    for n, (c, t) in agg.items():
        fh.write(f"| `{n}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |\n")
    fh.write(f"| **sum** | {sum(v[0] for v in agg.values())} | {tot:.1f} | 100% |\n")

# ---- full captures: key metrics per kernel
want = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def num(v, unit=""):
    x = float(v.replace(",", ""))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "us": 1e-3, "usecond": 1e-3, "ms": 1.0,
             "msecond": 1.0, "ns": 1e-6, "nsecond": 1e-6}.get(unit, 1.0)
    return x * scale


traffic_path = os.path.join(out_dir, "r02_ncu_kernels.json")
traffic = {}  # rebuilt from the captures given on the command line (never merged with an older capture)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hashlib  # noqa: E402


def kernel_source_hash():  # must equal bench.kernel_source_hash()
    cs = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "parallel-in-time-ode-filters_b200",
                      "csrc")
    h = hashlib.sha256()
    for f in ("pof_lane2.cuh", "pof_lane2_kernels.cuh", "pof_small.cuh"):
        h.update(open(os.path.join(cs, f), "rb").read())
    return h.hexdigest()[:16]


traffic["kernel_source_hash"] = kernel_source_hash()
seen = {}
with open(f"{out_dir}/{tag}_ncu_full_summary.md", "w") as fh:
    fh.write(f"# {tag}: ncu --set full --clock-control none, first launch of each kernel\n\n")
    for one in [rep] + more_reps:
        raw = subprocess.run(["ncu", "-i", one, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rr = list(csv.reader(raw.splitlines()))
        hdr, units = rr[0], rr[1]
        idx = {h: i for i, h in enumerate(hdr)}
        for r in rr[2:]:
            name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").replace("pof::", "").replace(" ", "")
            if name in seen:
                continue
            seen[name] = 1
            fh.write(f"## `{name}`\n\n| metric | value | unit |\n|---|---|---|\n")
            for w in want:
                if w in idx:
                    fh.write(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |\n")
            st = [(h, r[i]) for h, i in idx.items()
                  if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h]
            st = [(h, float(v.replace(",", ""))) for h, v in st if v not in ("", "n/a")]
            st.sort(key=lambda x: -x[1])
            fh.write("\nTop stall reasons (warps stalled per issue-active cycle): " + ", ".join(
                f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}={v:.2f}"
                for h, v in st[:6]) + "\n\n")
            g = lambda m: num(r[idx[m]], units[idx[m]])
            traffic[name] = {
                "dram_bytes_read": g("dram__bytes_read.sum"), "dram_bytes_write": g("dram__bytes_write.sum"),
                "dram_bytes": g("dram__bytes_read.sum") + g("dram__bytes_write.sum"),
                "duration_ms_under_ncu": g("gpu__time_duration.sum"),
                "fp64_pipe_pct": g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
                "lsu_pipe_pct": g("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "registers": g("launch__registers_per_thread"),
                "source": f"profiles/{tag}_ncu_full_summary.md (ncu --set full --clock-control none)",
            }
json.dump(traffic, open(traffic_path, "w"), indent=1)
print("wrote", f"{out_dir}/{tag}_launch_shares.md", f"{out_dir}/{tag}_ncu_full_summary.md", traffic_path)
