"""Whole-model training step (forward + backward), BASELINE config c2: RWKV-7 0.4B (D=1024, L=24, H=16, vocab 8193),
batch 8 x 4096 tokens, bf16, random-init weights, synthetic ids, one B200.  Reports tokens/s with the fused time-mix
kernels and with the ATen elementwise chain (same WKV kernels in both), i.e. what the fused kernels buy end to end.
usage: python scripts/bench_train_step.py [batch] [steps]"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rwkvtts_b200 import core
from rwkvfla.models.rwkv7 import RWKV7Config, RWKV7ForCausalLM

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T = 4096
torch.manual_seed(42)
cfg = RWKV7Config(hidden_size=1024, num_hidden_layers=24, head_dim=64, vocab_size=8193, decay_low_rank_dim=64,
                  a_low_rank_dim=64, v_low_rank_dim=32, gate_low_rank_dim=128, fuse_cross_entropy=True)
m = RWKV7ForCausalLM(cfg)
with torch.no_grad():
    for _, p in m.named_parameters():
        if p.abs().sum() == 0:
            p.copy_(torch.randn_like(p) * 0.02)
m = m.cuda().to(torch.bfloat16).train()
ids = torch.randint(0, 8192, (B, T), device="cuda")
labels = ids.clone()


def step():
    out = m(input_ids=ids, labels=labels)
    out.loss.backward()
    return out.loss


res = {}
for fused in (True, False):
    core.FUSED = fused
    for _ in range(2):
        m.zero_grad(set_to_none=True)
        loss = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        m.zero_grad(set_to_none=True)
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    res["fused" if fused else "aten"] = {"ms_per_step": ms, "tokens_per_s": B * T / ms * 1e3, "loss": float(loss),
                                        "peak_mem_GB": torch.cuda.max_memory_allocated() / 1e9}
core.FUSED = True
print(json.dumps({"metric": "whole-model train step fwd+bwd tokens/s, RWKV-7 0.4B, seq 4096, 1 GPU", "batch": B, **res}))
