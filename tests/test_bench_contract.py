"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints ONE JSON line with the
driver's keys (measured at the stated N, never extrapolated), the native arm refuses to run without a CUDA device
(no CPU fallback of the product path), and the roofline capture is tied to the built kernel sources by hash."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_line():
    p = _run("--impl", "reference", "--log2n", "10", "--steps", "2", "--warmup", "1")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in j, k
    assert j["impl"] == "reference" and j["higher_is_better"] is False and j["gpu_launches"] == 0
    assert j["config"]["n_time_total"] == 1024 and j["config"]["same_config"] is True
    assert j["value"] > 0 and j["e2e"]["value"] == j["value"] and j["e2e"]["h2d_bytes_per_step"] == 0
    assert j["cpu_baseline"]["kind"] == "port" and "MEASURED" in j["cpu_baseline"]["sample"]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour without a GPU")
def test_native_arm_needs_a_gpu():
    p = _run("--steps", "1", "--warmup", "0")
    assert p.returncode != 0
    assert "no CPU fallback" in (p.stderr + p.stdout)


def test_committed_ncu_capture_matches_the_built_kernels():
    sys.path.insert(0, ROOT)
    import bench

    t = json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_kernels.json")))
    assert t["kernel_source_hash"] == bench.kernel_source_hash(), \
        "profiles/r02_ncu_kernels.json was captured for different leaf-kernel sources: re-run scripts/gpu_profiles.sh"
    assert 0.3 < t["k_lane2_scan<2,3>"]["fp64_pipe_pct"] / 100.0 < 1.0
