"""bench.py host-side behaviour that needs no GPU: the reference arm's JSON line (contract keys, `extrapolated` label),
and the leg runner's failure handling (a leg that dies or hangs must not take the headline line with it)."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, RWKVTTS_BENCH_CPU_SECONDS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "3",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["extrapolated"] is True and "EXTRAPOLATED" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["metric"] == _bench().METRIC


def test_other_ranks_of_the_reference_arm_exit_without_work():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_a_failing_or_hanging_leg_is_recorded_not_fatal(monkeypatch):
    b = _bench()
    real = subprocess.run

    def fake_fail(cmd, **kw):
        return subprocess.CompletedProcess(cmd, 3, stdout="", stderr="boom")

    def fake_hang(cmd, **kw):
        raise subprocess.TimeoutExpired(cmd, kw.get("timeout"))
    monkeypatch.setattr(b.subprocess, "run", fake_fail)
    assert "error" in b.run_leg("decode", 5) and "boom" in b.run_leg("decode", 5)["stderr_tail"]
    monkeypatch.setattr(b.subprocess, "run", fake_hang)
    assert "killed" in b.run_leg("decode", 5)["error"]
    monkeypatch.setattr(b.subprocess, "run", real)


def test_every_leg_has_a_timeout_and_a_function():
    b = _bench()
    assert {n for n, _ in b.LEGS} == set(b.LEG_FN)
    assert all(10 <= t <= 600 for _, t in b.LEGS)
