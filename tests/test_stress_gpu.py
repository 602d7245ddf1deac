"""Reliability of the chunked tcgen05 pair (round-1 VERDICT: the backward shipped with a barrier hazard).

* thousands of back-to-back launches at the c2 and c5 shapes finish and stay bit-identical;
* with group C2 of the backward stalled for about a chunk (build variant `delay`) the shipped protocol -- one
  `out_ready` barrier per iteration parity -- still finishes;
* the round-1 protocol under the same stall (variant `oldbar`) does not: the watchdog turns the deadlock into a CUDA
  error and names the barrier.  Each run is its own process: a tripped watchdog loses the CUDA context.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "scripts", "stress_wkv7.py")


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, SCRIPT, *args], capture_output=True, text=True, timeout=timeout)


def _have(variant):
    return os.path.exists(os.path.join(ROOT, "rwkvtts_b200", f"librwkvtts_wkv7_{variant}.so"))


@pytest.mark.gpu
@pytest.mark.parametrize("shape,pairs", [("c2", 5000), ("c5", 3000), ("small", 3000)])
def test_back_to_back_launches_finish_and_are_bit_identical(shape, pairs):
    r = _run("--shape", shape, "--pairs", str(pairs))
    assert r.returncode == 0, r.stdout + r.stderr
    assert "stress ok" in r.stdout


@pytest.mark.gpu
def test_shipped_protocol_survives_a_stalled_output_group():
    if not _have("delay"):
        pytest.skip("variant not built (python -m rwkvtts_b200.build --variants)")
    r = _run("--variant", "delay", "--shape", "c2", "--pairs", "300", "--check-every", "100")
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_round1_single_barrier_deadlocks_under_the_same_stall_and_the_watchdog_names_it():
    if not _have("oldbar"):
        pytest.skip("variant not built (python -m rwkvtts_b200.build --variants)")
    r = _run("--variant", "oldbar", "--shape", "c2", "--pairs", "300", "--check-every", "100", timeout=120)
    assert r.returncode == 3, r.stdout + r.stderr
    assert "out_ready" in r.stdout, r.stdout + r.stderr
