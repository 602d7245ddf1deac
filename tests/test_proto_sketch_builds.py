"""The forward-v2 sketch (proto/wkv7_tc_fwd_v2.cu) is not part of the library and has never run; this
only keeps it compiling against the library's headers for sm_100a (nvcc cross-compiles without a GPU), so that it is a
usable starting point."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_forward_v2_sketch_compiles(tmp_path):
    out = subprocess.run(["nvcc", "-c", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xptxas", "-v",
                          "-I", os.path.join(ROOT, "rwkvtts_b200", "csrc"), os.path.join(ROOT, "proto", "wkv7_tc_fwd_v2.cu"),
                          "-o", str(tmp_path / "fwd_v2.o")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    log = out.stdout + out.stderr
    assert log.count("Compiling entry function") == 2          # inference and training variants
    assert "bytes spill stores" in log and " 0 bytes spill stores" in log
