"""x070.RWKV7S2S_SingleFFN against the reference's own module
(model/llm/rwkv_s2s_single_ffn.py:61-330, imported as tests/golden/make_golden.py imports it: deepspeed and the JIT
build of the CUDA ops stubbed) on CPU, forward AND backward, with the WKV op bound to the differentiable f64 oracle on
both sides.  Pins everything around the op -- embedding, masks, token shift, lerps, LoRAs, kk / k update, GroupNorm,
bonus, gate, FFN, residuals, ln_out, the two heads -- and its gradients; the op itself is pinned on the GPU
(tests/test_wkv7_gpu.py).  Build container only."""
import os
import sys
from argparse import Namespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.exists("/root/reference/model/llm/rwkv_s2s_single_ffn.py"),
                                reason="reference tree not mounted")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def _oracle_op(r, w, k, v, a, b):
    """RUN_CUDA_RWKV7g (:37-40): [B,T,H*64] tensors in the BlinkDL order; bf16 I/O like the CUDA op, differentiable."""
    from oracle.wkv7_oracle import wkv7_forward
    B, T, HC = r.shape
    q = lambda t: t.to(torch.bfloat16).double().view(B, T, HC // 64, 64) + (t - t.detach()).double().view(B, T, HC // 64, 64)
    y = wkv7_forward(q(w), q(r), q(k), q(v), q(a), q(b))
    y = y[0] if isinstance(y, tuple) else y
    y = y + (y.to(torch.bfloat16).double() - y).detach()          # bf16 output, straight-through gradient
    return y.reshape(B, T, HC).to(r.dtype)


def test_model_forward_and_backward_match_reference_module(monkeypatch):
    import make_golden
    from rwkvtts_b200 import core, x070
    import torch.utils.cpp_extension as ce
    saved_ds, saved_load = sys.modules.get("deepspeed"), ce.load
    ref = make_golden.import_reference()
    ce.load = saved_load                                  # import_reference() stubs the JIT build for the import only
    if saved_ds is not None:
        sys.modules["deepspeed"] = saved_ds                # import_reference() installs a stub
    else:
        sys.modules.pop("deepspeed", None)
    monkeypatch.setattr(ref, "RUN_CUDA_RWKV7g", _oracle_op)
    monkeypatch.setattr(core, "_wkv", lambda r, w, k, v, a, b, state, need_state, inplace_state=False:
                        (_oracle_op(r.float(), w.float(), k.float(), v.float(), a.float(), b.float()).to(r.dtype), None))
    args = Namespace(n_embd=128, n_layer=2, head_size_a=64, head_size_divisor=8, dim_att=128, dim_ffn=512, dropout=0.0,
                     vocab_size=97, text_vocab_size=53, audio_vocab_size=41, grad_cp=0, need_init_tmix=True, need_init_cmix=True)
    torch.manual_seed(0)
    rm = ref.RWKV7S2S_SingleFFN(args)
    with torch.no_grad():
        for p in rm.parameters():                          # the reference zero-inits output / value / LoRA-A weights
            if float(p.abs().sum()) == 0:
                p.normal_(0, 0.05)
    mm = x070.RWKV7S2S_SingleFFN(args)
    res = mm.load_state_dict(rm.state_dict(), strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    B, T = 2, 16
    idx = torch.randint(0, 97, (B, T))
    mask = torch.ones(B, T, dtype=torch.bool)
    mask[1, 11:] = False
    for is_text in (True, False):
        rm.zero_grad(); mm.zero_grad()
        lr_ = rm(idx, mask, is_text=is_text)
        lm_ = mm(idx, mask, is_text=is_text)
        a, b = (lr_[0], lm_[0]) if is_text else (lr_[1], lm_[1])
        assert (lr_[1] is None) == (lm_[1] is None) and (lr_[0] is None) == (lm_[0] is None)
        assert a.shape == b.shape == (B, T, 53 if is_text else 41)
        assert float((a.detach() - b.detach()).norm() / a.detach().norm()) < 1e-5
        w = torch.randn_like(a)
        (a * w).sum().backward()
        (b * w).sum().backward()
        gr, gm = dict(rm.named_parameters()), dict(mm.named_parameters())
        worst = max((float((gr[n].grad - gm[n].grad).norm() / gr[n].grad.norm().clamp(min=1e-12)), n)
                    for n in gr if gr[n].grad is not None)
        assert all((gr[n].grad is None) == (gm[n].grad is None) for n in gr)
        assert worst[0] < 1e-3, worst        # this repo's path hands the op bf16 copies, so the op's input gradients pass through a bf16 cast

