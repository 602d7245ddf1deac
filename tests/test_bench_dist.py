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
bench.dist_barrier(w)
m = bench.dist_max(10.0 + r, w)
bench.dist_barrier(w)
import torch.distributed as dist
assert dist.get_world_size() == w
print("rank", r, "max", m, flush=True)
dist.destroy_process_group()
"""


def test_bench_control_plane_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29577", str(script)],
                         capture_output=True, text=True, timeout=180, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("rank")]
    assert sorted(lines) == ["rank 0 max 11.0", "rank 1 max 11.0"], out.stdout


def test_bench_single_process_helpers_are_noops():
    sys.path.insert(0, ROOT)
    import bench
    bench.dist_init(1)
    assert bench.dist_max(3.5, 1) == 3.5
