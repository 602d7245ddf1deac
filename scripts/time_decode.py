"""Times the decode step of the 0.4B model at batch 32 (BASELINE configs[3]): one-kernel step (csrc/decode_step.cu) and
the CUDA-graph step it replaces, CUDA events over back-to-back steps.  Usage: python scripts/time_decode.py [steps] [batch]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from rwkvfla.models.rwkv7 import Cache
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
if "--no-graph" not in sys.argv:
    with torch.no_grad():
        g = _GraphDecodeStep(m, cache, B, dev)
        t_graph = timed(lambda: g(tok), steps)
    print(f"CUDA-graph step, B={B}: {t_graph:.3f} ms/step = {B / t_graph * 1e3:.0f} tokens/s")
