"""bench.py's multi-GPU control plane (barrier, max over ranks) with world size 2 over gloo on CPU: the N > 1 path of
the bench has no data-path collective, these two helpers are all the ranks exchange."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import bench
w, r = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"])
bench.dist_init(w)
assert bench._CTL["mode"] == os.environ["EXPECT_MODE"], bench._CTL
bench.dist_barrier(w)
m = bench.dist_max(10.0 + r, w)
bench.dist_barrier(w)
m2 = bench.dist_max(-1.0 - r, w)
if bench._CTL["mode"] == "dist":
    import torch.distributed as dist
    assert dist.get_world_size() == w
sys.stdout.write("rank %d max %s %s\n" % (r, m, m2)); sys.stdout.flush()     # one write per rank: lines do not interleave
bench.dist_finish(w)
"""


def _run_world2(tmp_path, port, extra_env, mode):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", EXPECT_MODE=mode, **extra_env)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(script)],
                         capture_output=True, text=True, timeout=180, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("rank")]
    assert sorted(lines) == ["rank 0 max 11.0 -1.0", "rank 1 max 11.0 -1.0"], out.stdout


def test_bench_control_plane_world_size_2(tmp_path):
    _run_world2(tmp_path, 29577, {}, "dist")


def test_bench_control_plane_file_fallback_world_size_2(tmp_path):
    """the fallback the bench takes when no process group can be created: same barrier / max over files in /tmp"""
    _run_world2(tmp_path, 29578, {"RWKVTTS_BENCH_CONTROL": "fs"}, "fs")


def test_bench_single_process_helpers_are_noops():
    sys.path.insert(0, ROOT)
    import bench
    bench.dist_init(1)
    assert bench.dist_max(3.5, 1) == 3.5


def test_reference_arm_line_has_the_contract_keys(capsys):
    """`bench.py --impl reference`: the reference algorithm (C port of the reference kernels) on the host cores, bounded
    sample; the line carries the contract keys (impl, metric, value, unit, cpu_baseline{kind, cores, sample}, e2e{...})."""
    import json
    import types
    sys.path.insert(0, ROOT)
    import bench
    cb = bench.cpu_reference(seconds_target=0.5)
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "T=4096" in cb["sample"]
    orig = bench.cpu_reference
    bench.cpu_reference = lambda seconds_target=20.0: cb          # keep the test short: reuse the measured sample
    try:
        bench.run_reference(types.SimpleNamespace(gpus=2, steps=2, warmup=1), rank=1)     # other ranks: no work, no line
        assert capsys.readouterr().out == ""
        bench.run_reference(types.SimpleNamespace(gpus=2, steps=2, warmup=1), rank=0)
    finally:
        bench.cpu_reference = orig
    line = json.loads(capsys.readouterr().out)
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == "tokens/s"
    assert line["higher_is_better"] is True and line["value"] == line["cpu_baseline"]["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"] == bench.WORKLOAD and line["n_gpus"] == 2
