"""Synthetic op-level inputs for the WKV-7 path, exactly as SURVEY.md section 8(d) / BASELINE.md 3.2
specify them (seed 42 is the reference scripts' default, train_spark_rwkv7speech.py:107-110)."""
import torch


def make_inputs(B, T, H, C=64, seed=42, dtype=torch.bfloat16, k_update=True):
    """r,k,v,dy ~ N(0,1); w = -softplus(-N(0,1)) - 0.5 (decay exp(-exp(w)) in (0.545,1));
    kk = normalize(N(0,1)); g = sigmoid(N(0,1)); kernel args a = -kk, b = kk*g;
    optional k <- k*(1+(g-1)*1.02) (rwkv_s2s_single_ffn.py:189).
    Returns a dict of contiguous CPU tensors [B,T,H,C] in `dtype`: w,q,k,v,a,b,dy."""
    gen = torch.Generator().manual_seed(seed)
    n = lambda: torch.randn(B, T, H, C, generator=gen, dtype=torch.float32)
    q, k, v, xw, kk, g, dy = n(), n(), n(), n(), n(), n(), n()
    w = -torch.nn.functional.softplus(-xw) - 0.5
    kk = torch.nn.functional.normalize(kk, dim=-1, p=2.0)
    g = torch.sigmoid(g)
    if k_update:
        k = k * (1 + (g - 1) * 1.02)
    out = dict(w=w, q=q, k=k, v=v, a=-kk, b=kk * g, dy=dy)
    return {n_: x.to(dtype).contiguous() for n_, x in out.items()}
