"""The C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol include/pof_b200.h declares.
No compute calls here (no GPU in this container)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pof_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pof_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(native_lib):
    lib = ctypes.CDLL(native_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pof_b200.h but not exported"
    assert sorted(native_lib.EXPORTED) == names


def test_host_only_queries(native_lib):
    L = native_lib.LIB
    assert L.pof_supported(2, 3) == 1 and L.pof_supported(1, 1) == 1 and L.pof_supported(4, 5) == 1
    # beyond the (d <= 4)-templated families: the large-state tile kernels, limited by shared memory (D <= 64 at d = 16)
    assert L.pof_supported(5, 3) == 1 and L.pof_supported(16, 3) == 1 and L.pof_supported_tile(2, 3) == 1
    assert L.pof_supported(16, 4) == 0 and L.pof_supported(2, 9) == 0 and L.pof_supported_tile(32, 3) == 0
    assert L.pof_default_chunk_len(1 << 18, 16, 3, 148, 0) == ((1 << 18) - 1 + 147) // 148  # one chunk per SM
    assert L.pof_default_chunk_len(1 << 20, 2, 3, 148, 0) >= 4
    nb = L.pof_workspace_bytes(1 << 20, 2, 3, 222)
    assert 1.1e9 < nb < 1.4e9  # dominated by the per-step backward kernels: n * 136 doubles
    # dataflow tree sweeps: 3 leaf kernels + chunk elements + 3 tree launches + 2 reductions; one launch per tree level
    # (flag POF_F_TREE_PER_LEVEL, kept for A/B measurements) needs ~6x as many
    assert L.pof_launches_per_pass(1 << 20, 2, 3, 222, 0) == 9 and L.pof_launches_per_pass(1 << 14, 2, 3, 4, 0) == 9
    assert L.pof_launches_per_pass(1 << 20, 2, 3, 222, 4) > 40
    # the forced large-state family fills the GPU with one chunk per resident CTA
    assert L.pof_default_chunk_len(1 << 20, 2, 3, 148, 1) > L.pof_default_chunk_len(1 << 20, 2, 3, 148, 0)


def test_sass_is_sm100a(native_lib):
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", native_lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_library_reads_no_environment(native_lib):
    """kernel-family choices are explicit `flags` arguments of the C ABI: the library imports no getenv"""
    csrc = os.path.join(ROOT, "parallel-in-time-ode-filters_b200", "csrc")
    for f in sorted(os.listdir(csrc)):  # (the statically linked CUDA runtime imports getenv itself: check OUR sources)
        assert "getenv" not in open(os.path.join(csrc, f)).read(), f


def test_no_cpu_fallback(monkeypatch):
    """the product path refuses CPU tensors"""
    import torch
    from pof import _native as nat

    with pytest.raises(nat.NativeError):
        nat.require_cuda(torch.zeros(3, dtype=torch.float64))


def test_header_is_plain_c_and_links(native_lib, tmp_path):
    """The boundary is a C ABI: include/pof_b200.h compiles as ISO C99 (no C++ or torch types in any signature) and a
    plain C program links against libpof_b200.so and calls its host-only queries."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include "pof_b200.h"\n'
        "int main(void) {\n"
        "  pof_ctx_t* ctx = 0; pof_loop_t* loop = 0; pof_p2p_t* p2p = 0;\n"
        "  size_t b = pof_workspace_bytes(1 << 20, 2, 3, 111);\n"
        "  size_t b32 = pof_workspace_bytes_f32(1 << 20, 2, 3, 111);\n"
        "  long long launches = (long long)pof_launches_per_pass(1 << 20, 2, 3, 111, 0u);\n"
        "  (void)ctx; (void)loop; (void)p2p;\n"
        '  printf("%d %d %lld %d\\n", pof_supported(2, 3), b > b32 && b32 > 0, launches, (int)POF_F_TREE_UPDOWN);\n'
        "  return 0;\n}\n")
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(native_lib.LIB_PATH)
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, str(src), "-o", str(exe),
                    "-L", libdir, "-lpof_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["1", "1", "9", "16"]
