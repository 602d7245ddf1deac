"""Generate golden vectors by running the REFERENCE's own Python on the CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference module ``model/llm/rwkv_s2s_single_ffn.py`` imports ``deepspeed`` and
JIT-builds its CUDA ops at import time (:9-13, :42-43); both are stubbed here so that the
module's plain-tensor code can run on the CPU.  Nothing of the reference is modified or
copied: the functions are imported from where they lie and called.

Fixtures written (small, committed):

* ``tmix_one_L{0,1}.pt``  -- ``RWKV_x070_TMix_one`` (:482-506) iterated over T steps:
  inputs, weights, per-step outputs, final state.  This is the executable spec of one decode
  step of time-mix *including the WKV state update* (:497-502) and so pins the oracle's
  recurrence, its ``w`` convention and the value-major state layout.
* ``cmix_one.pt``         -- ``RWKV_x070_CMix_one`` (:545-549).
* ``block_L{0,1}.pt``     -- training-time ``Block.forward`` (:251-259) = ``RWKV_Tmix_x070``
  (:158-196) + ``RWKV_CMix_x070`` (:223-230) with the CUDA op ``RUN_CUDA_RWKV7g`` replaced
  by the reference's own "cuda-free method" loop (:526-533) in fp32 -- pins everything
  AROUND the WKV op (token shift, lerps, LoRAs, kk/k update, GroupNorm, bonus, gate, FFN).
"""
import importlib.util
import os
import sys
import types
from argparse import Namespace

import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    ds = types.ModuleType("deepspeed")
    ds.checkpointing = types.SimpleNamespace(checkpoint=lambda fn, *a: fn(*a))
    sys.modules["deepspeed"] = ds
    import torch.utils.cpp_extension as ce
    ce.load = lambda *a, **k: None          # no GPU here: skip the JIT build of the CUDA ops
    spec = importlib.util.spec_from_file_location(
        "ref_rwkv_s2s_single_ffn", f"{REF}/model/llm/rwkv_s2s_single_ffn.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cuda_free_wkv(q, w, k, v, a, b):
    """The reference's "cuda-free method" (:526-533) as a stand-in for RUN_CUDA_RWKV7g
    (argument order of :37: q,w,k,v,a,b with a=-kk, b=kk*a; w = BlinkDL pre-activation)."""
    B, T, HC = q.shape
    H, N = HC // 64, 64
    state = torch.zeros(B, H, N, N, dtype=torch.float32)
    out = torch.empty(B, T, HC, dtype=torch.float32)
    dec = torch.exp(-torch.exp(w.float()))
    for t in range(T):
        r_, w_, k_, v_, a_, b_ = [x[:, t].float().view(B, H, N) for x in (q, dec, k, v, a, b)]
        vk = v_.unsqueeze(-1) @ k_.unsqueeze(-2)
        ab = a_.unsqueeze(-1) @ b_.unsqueeze(-2)
        state = state * w_.unsqueeze(-2) + state @ ab + vk
        out[:, t] = (state @ r_.unsqueeze(-1)).view(B, HC)
    return out.to(q.dtype)


def main():
    ref = import_reference()
    torch.manual_seed(42)
    H, N = 2, 64
    C = H * N
    T = 24

    # ---- TMix_one / CMix_one (decode-step spec), fp32 weights --------------------------
    def rnd(*s, scale=1.0):
        return torch.randn(*s) * scale

    for layer_id in (0, 1):
        wts = dict(
            x_r=torch.rand(C), x_w=torch.rand(C), x_k=torch.rand(C), x_v=torch.rand(C),
            x_a=torch.rand(C), x_g=torch.rand(C),
            w0=rnd(C), w1=rnd(C, 32, scale=0.1), w2=rnd(32, C, scale=0.3),
            a0=rnd(C, scale=0.5), a1=rnd(C, 32, scale=0.1), a2=rnd(32, C, scale=0.3),
            v0=rnd(C, scale=0.5), v1=rnd(C, 32, scale=0.1), v2=rnd(32, C, scale=0.3),
            g1=rnd(C, 32, scale=0.1), g2=rnd(32, C, scale=0.3),
            k_k=torch.full((C,), 0.71) + rnd(C, scale=0.05), k_a=torch.full((C,), 1.02),
            r_k=rnd(C, scale=0.1),
            R_=rnd(C, C, scale=C ** -0.5), K_=rnd(C, C, scale=C ** -0.5),
            V_=rnd(C, C, scale=C ** -0.5), O_=rnd(C, C, scale=C ** -0.5),
            ln_w=1 + rnd(C, scale=0.1), ln_b=rnd(C, scale=0.1),
        )
        xs = rnd(T, C)
        vf_in = rnd(T, C)
        x_prev = rnd(C, scale=0.5)
        state = rnd(H, N, N, scale=0.2)
        rec = dict(layer_id=layer_id, H=H, N=N, weights=wts, x=xs, v_first_in=vf_in,
                   x_prev0=x_prev.clone(), state0=state.clone())
        outs, vfs = [], []
        order = ["x_r", "x_w", "x_k", "x_v", "x_a", "x_g", "w0", "w1", "w2", "a0", "a1", "a2",
                 "v0", "v1", "v2", "g1", "g2", "k_k", "k_a", "r_k", "R_", "K_", "V_", "O_",
                 "ln_w", "ln_b"]
        for t in range(T):
            o, x_prev, state, vf = ref.RWKV_x070_TMix_one(
                layer_id, H, N, xs[t], x_prev, vf_in[t], state, *[wts[n] for n in order])
            outs.append(o)
            vfs.append(vf)
        rec.update(out=torch.stack(outs), v_first_out=torch.stack(vfs), x_prev_T=x_prev,
                   state_T=state)
        torch.save(rec, f"{OUT}/tmix_one_L{layer_id}.pt")
        print("tmix_one", layer_id, rec["out"].abs().mean().item(), state.abs().mean().item())

    xk, K_, V_ = torch.rand(C), rnd(C, 4 * C, scale=C ** -0.5), rnd(4 * C, C, scale=(4 * C) ** -0.5)
    xs, x_prev = rnd(T, C), rnd(C)
    rec = dict(x=xs, x_prev0=x_prev.clone(), x_k=xk, K_=K_, V_=V_)
    outs = []
    for t in range(T):
        o, x_prev = ref.RWKV_x070_CMix_one(xs[t], x_prev, xk, K_, V_)
        outs.append(o)
    rec.update(out=torch.stack(outs), x_prev_T=x_prev)
    torch.save(rec, f"{OUT}/cmix_one.pt")

    # ---- training-time Block.forward around a stand-in WKV ------------------------------
    ref.RUN_CUDA_RWKV7g = cuda_free_wkv
    args = Namespace(n_layer=2, n_embd=C, head_size_a=N, head_size_divisor=8, dropout=0.0,
                     need_init_tmix=True, need_init_cmix=True)
    B, T2 = 2, 32
    for layer_id in (0, 1):
        blk = ref.Block(args, layer_id).float()
        with torch.no_grad():     # the reference zero-inits these (:113,:117,:156,:221): make them live
            for n, p in blk.named_parameters():
                if p.abs().sum() == 0:
                    p.copy_(torch.randn_like(p) * 0.1)
        x = rnd(B, T2, C)
        mask = torch.ones(B, T2, 1)
        mask[1, :5] = 0           # left padding on sample 1 (spark_dataset.py:163-239)
        v_first = rnd(B, T2, C)
        with torch.no_grad():
            y, vf = blk(x, mask, v_first.clone())
        torch.save(dict(layer_id=layer_id, args=vars(args), state_dict=blk.state_dict(), x=x,
                        mask=mask, v_first_in=v_first, y=y, v_first_out=vf),
                   f"{OUT}/block_L{layer_id}.pt")
        print("block", layer_id, y.abs().mean().item())


if __name__ == "__main__":
    main()
