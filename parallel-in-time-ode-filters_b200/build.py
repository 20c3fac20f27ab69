"""Builds libpof_b200.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

    python parallel-in-time-ode-filters_b200/build.py [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libpof_b200.so")
SOURCES = [
    "pof_api.cu", "pof_seq_d1.cu", "pof_seq_d2.cu", "pof_seq_d3.cu", "pof_seq_d4.cu",
    "pof_tree_a.cu", "pof_tree_b.cu", "pof_tree_c.cu",
    "pof_lane2_d1.cu", "pof_lane2_d2.cu", "pof_lane2_d3.cu", "pof_lane2_d4.cu",
    "pof_tile.cu",
    # the optional fp32 mode: the same register-resident sources compiled with the scalar type float (pof_real.cuh)
    "pof_api_f32.cu", "pof_lane2_d1_f32.cu", "pof_lane2_d2_f32.cu", "pof_lane2_d3_f32.cu", "pof_lane2_d4_f32.cu",
    "pof_tree_a_f32.cu", "pof_tree_b_f32.cu", "pof_tree_c_f32.cu",
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
] + os.environ.get("POF_NVCC_EXTRA", "").split()  # e.g. -DPOF_TUNE: tuning hooks of scripts/tune_tree.py (never shipped)


HASH = LIB + ".srchash"


def _source_digest():
    import hashlib

    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in sorted(_headers() + [os.path.join(CSRC, s) for s in SOURCES]):
        h.update(os.path.basename(path).encode())
        h.update(open(path, "rb").read())
    return h.hexdigest()


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def _headers():
    inc = os.path.join(os.path.dirname(HERE), "include")
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [
        os.path.join(inc, f) for f in os.listdir(inc)
    ]


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdr_time = _newest(_headers())
    # up to date: the library was built from exactly these sources (content hash in a sidecar file; file times do not
    # survive the copy to the GPU box reliably, and the object directory does not travel -- the library does)
    digest = _source_digest()
    if not force and os.path.exists(LIB) and os.path.exists(HASH) and open(HASH).read().strip() == digest:
        if verbose:
            print("up to date:", LIB)
        return LIB
    # object files are reused only if they were compiled with exactly these flags (a sidecar file per build directory)
    # and are newer than every source / header; otherwise everything is recompiled
    flags_file = os.path.join(OBJ, "nvcc_flags.txt")
    flags_now = " ".join(NVCC_FLAGS)
    flags_same = os.path.exists(flags_file) and open(flags_file).read().strip() == flags_now
    jobs = []
    objs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(op)
        dep_time = max(os.path.getmtime(sp), hdr_time)
        if src.endswith("_f32.cu"):  # the fp32 translation units #include their fp64 twin
            dep_time = max(dep_time, os.path.getmtime(os.path.join(CSRC, src.replace("_f32.cu", ".cu"))))
        if force or not flags_same or not os.path.exists(op) or os.path.getmtime(op) < dep_time:
            jobs.append((sp, op))

    def run(job):
        sp, op = job
        cmd = [nvcc] + NVCC_FLAGS + ["-c", sp, "-o", op]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(op + ".log", "w") as fh:
            fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {sp}:\n{r.stdout}\n{r.stderr}")
        return op

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(run, jobs))
    with open(flags_file, "w") as fh:
        fh.write(flags_now + "\n")
    # the digest differs from the library's (or there is no library): always relink
    if True:
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(HASH, "w") as fh:
        fh.write(digest + "\n")
    if verbose:
        print("built", LIB, "(recompiled %d objects)" % len(jobs))
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
