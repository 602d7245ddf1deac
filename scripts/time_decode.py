"""Times the decode step of the 0.4B model at batch 32 (BASELINE configs[3]): one-kernel step (csrc/decode_step.cu) and
the CUDA-graph step it replaces, CUDA events over back-to-back steps.  Usage: python scripts/time_decode.py [steps] [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rwkvfla.models.utils import Cache
from rwkvfla.models.rwkv7.modeling_rwkv7 import _GraphDecodeStep
from rwkvtts_b200.decode import MegaDecodeStep

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device("cuda", 0)
m = bench.random_init_0p4b().to(dev).eval()
torch.manual_seed(7)
ids = torch.randint(0, 8192, (B, 163), device=dev)
with torch.no_grad():
    out = m(input_ids=ids, past_key_values=Cache(), use_cache=True, logits_to_keep=1)
cache = out.past_key_values
tok = out.logits[:, -1].float().argmax(-1)


def timed(fn, n):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


mega = MegaDecodeStep(m, cache, B, dev)
t_logits = timed(lambda: mega(tok), steps)
t0 = time.perf_counter()
toks = mega.greedy(tok, steps)
torch.cuda.synchronize()
t_greedy = (time.perf_counter() - t0) / steps * 1e3
print(f"one-kernel step, B={B}: {t_logits:.3f} ms/step (logits), {t_greedy:.3f} ms/step (greedy on device, wall) "
      f"= {B / t_greedy * 1e3:.0f} tokens/s")
pt = torch.stack([mega.phase_times(tok) for _ in range(20)])[5:].mean(0).cpu()
L = len(m.model.layers)
per = pt[: 7 * L].view(L, 7)
names = ["ln1", "proj", "wkv", "out", "ln2", "key", "value"]
print("phase times, us (mean over layers 1..L-1 | layer 0): " + ", ".join(
    f"{n} {per[1:, i].mean():.2f}|{per[0, i]:.2f}" for i, n in enumerate(names)) + f"; final norm {pt[7 * L]:.2f}; sum {pt.sum():.1f}")
# barrier anatomy of the last profiled step: per barrier, spread of arrivals and the time from the last arrival to the releases
arr, rel = mega.last_arrive.double().cpu(), mega.last_release.double().cpu()
G = int((arr[0] > 0).sum())
arr, rel = arr[:, :G], rel[:, :G]
last = arr.max(1).values
lat = (rel - last[:, None]) / 1e3                       # us from the last arrival to each CTA's release
work = (arr[1:] - rel[:-1]) / 1e3                       # us each CTA worked between two barriers
idx = lambda i: slice(7 + i, 7 * L, 7)                  # phase i of layers 1..L-1 (barrier index = 7 l + i)
print(f"grid {G}; barrier latency after the last arrival, us: median {lat.median():.2f}, max-over-CTAs mean {lat.max(1).values.mean():.2f}")
for i, n in enumerate(names):
    w = work[[k - 1 for k in range(7 + i, 7 * L, 7)]]
    print(f"  {n:6s} work per CTA, us: min {w.min(1).values.mean():.2f} median {w.median(1).values.mean():.2f} max {w.max(1).values.mean():.2f}"
          f"   barrier latency (max over CTAs) {lat[idx(i)].max(1).values.mean():.2f}")
f = mega.last_fine.cpu().tolist()
d = lambda a, b: (f[b] - f[a]) if f[a] and f[b] else None
print("cycles, CTA 0, layer 1: prefetch_gemm(out) %s, prefetch_vecs %s, prefetch_wkv %s, stage_up %s" % (d(0, 1), d(1, 2), d(3, 4), d(5, 6)))
for name, base, per, n in (("proj", 16, 4, 4), ("wkv", 32, 8, 2), ("out", 48, 4, 2), ("key", 64, 4, 4)):
    for r in range(n):
        o = base + r * per
        if f[o + (1 if per == 4 else 0)]:
            pts = [f[o + k] for k in range(per + (1 if per == 4 else 0))]
            print(f"  {name} round {r}: " + " ".join(str(b - a) if a and b else "-" for a, b in zip(pts[:-1], pts[1:])))
for name, o in (("ln1", 80), ("ln2", 88)):
    if f[o] and f[o + 4]:
        print(f"  {name}: loads+row {f[o+1]-f[o]}, store residual {f[o+2]-f[o+1]}, LayerNorm {f[o+3]-f[o+2]}, shift/lerp/stores {f[o+4]-f[o+3]}")
if "--no-graph" not in sys.argv:
    with torch.no_grad():
        g = _GraphDecodeStep(m, cache, B, dev)
        t_graph = timed(lambda: g(tok), steps)
    print(f"CUDA-graph step, B={B}: {t_graph:.3f} ms/step = {B / t_graph * 1e3:.0f} tokens/s")
