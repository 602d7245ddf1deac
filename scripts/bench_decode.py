"""AR decode throughput, BASELINE config c4: RWKV-7 0.4B (D=1024, L=24, H=16, vocab 8193 Spark audio head), 32 prompts of
163 positions, greedy, EOS suppressed.  tokens/s = 32 * new_tokens / time after the prefill.
Random-init weights of that architecture (no checkpoints offline).  Prints one JSON line per mode."""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rwkvfla.models.rwkv7 import RWKV7Config, RWKV7ForCausalLM

B, PROMPT, NEW = 32, 163, int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.manual_seed(42)
cfg = RWKV7Config(hidden_size=1024, num_hidden_layers=24, head_dim=64, vocab_size=8193, decay_low_rank_dim=64,
                  a_low_rank_dim=64, v_low_rank_dim=32, gate_low_rank_dim=128)
m = RWKV7ForCausalLM(cfg)
with torch.no_grad():
    for n, p in m.named_parameters():
        if p.abs().sum() == 0:
            p.copy_(torch.randn_like(p) * 0.02)
m = m.cuda().to(torch.bfloat16).eval()
ids = torch.randint(0, 8192, (B, PROMPT), device="cuda")
res = {}
for mode in ("eager", "cuda_graph"):
    kw = dict(input_ids=ids, do_sample=False, eos_token_id=None, use_cuda_graph=(mode == "cuda_graph"))
    m.generate(max_new_tokens=12, **kw)                      # warm-up (allocator, cuBLAS)
    torch.cuda.synchronize()
    SHORT = 16                                               # both runs pay prefill (+ graph capture): the difference is pure steps
    t0 = time.perf_counter(); m.generate(max_new_tokens=SHORT, **kw); torch.cuda.synchronize(); t_short = time.perf_counter() - t0
    t0 = time.perf_counter(); out = m.generate(max_new_tokens=NEW, **kw); torch.cuda.synchronize(); t_all = time.perf_counter() - t0
    res[mode] = out
    dt = t_all - t_short
    print(json.dumps({"metric": "AR decode tokens/s, RWKV-7 0.4B, batch 32, greedy", "mode": mode, "new_tokens": NEW,
                      "value": B * (NEW - SHORT) / dt, "ms_per_step": dt / (NEW - SHORT) * 1e3,
                      "prefill_plus_setup_ms": (t_short - SHORT * dt / (NEW - SHORT)) * 1e3}))
print("greedy ids identical:", bool(torch.equal(res["eager"], res["cuda_graph"])))
