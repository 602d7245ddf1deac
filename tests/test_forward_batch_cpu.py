"""x070.RWKV_Tmix_x070 / RWKV_CMix_x070 / Block `forward_batch` against the reference's own
classes (model/llm/rwkv_asr_cuda_whisper.py:98-326) on CPU.  The reference file JIT-builds its CUDA ops at import, so
its three class definitions are extracted with ast and executed with RWKV7_BATCH_OP bound to the f64 oracle of the
stateful op; this repo's modules run the same test with `core._wkv` bound to the same oracle (the CUDA op itself is
pinned by tests/test_wkv7_gpu.py).  What is compared is everything around the op: masking of x and v only, token-shift
states, in-place recurrent state, and the reference's return convention.  Build container only."""
import ast
import math
import os
import sys
from argparse import Namespace

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/model/llm/rwkv_asr_cuda_whisper.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted")
sys.path.insert(0, ROOT)


def _oracle_batch_op(state, r, w, k, v, a, b):
    from oracle.wkv7_oracle import wkv7_state_forward
    # the op's I/O is bf16 on both sides (the modules run in fp32 here so that everything else compares tightly)
    y, sT = wkv7_state_forward(state.double(), *(t.to(torch.bfloat16).double() for t in (r, w, k, v, a, b)))
    state.copy_(sT.to(state.dtype))                       # rwkv7_state_fwd_fp16.cu:54-56: in place
    return y.to(torch.bfloat16).to(r.dtype)


def _reference_classes():
    tree = ast.parse(open(REF).read())
    want = ("RWKV_Tmix_x070", "RWKV_CMix_x070", "Block")
    body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in want]
    assert len(body) == 3
    ns = {"torch": torch, "nn": nn, "F": F, "math": math, "HEAD_SIZE": 64, "RWKV7_BATCH_OP": _oracle_batch_op,
          "RUN_CUDA_RWKV7g": None, "deepspeed": None}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


def test_forward_batch_matches_reference_classes(monkeypatch):
    from rwkvtts_b200 import core, x070
    ref = _reference_classes()

    def wkv(r, w, k, v, a, b, state, need_state, inplace_state=False):
        st = state if state is not None else torch.zeros(r.shape[0], r.shape[2] // 64, 64, 64)
        st = st if inplace_state else st.clone()
        return _oracle_batch_op(st, r, w, k, v, a, b), st

    monkeypatch.setattr(core, "_wkv", wkv)
    args = Namespace(n_embd=128, n_layer=3, head_size_a=64, head_size_divisor=8, dim_att=128, dim_ffn=512, dropout=0.0,
                     need_init_tmix=True, need_init_cmix=True)
    torch.manual_seed(0)
    B, T, C = 2, 7, 128
    for layer_id in (0, 1):
        rb = ref["Block"](args, layer_id)
        with torch.no_grad():
            for p in rb.parameters():                      # the reference zero-inits several projections: make them count
                if float(p.abs().sum()) == 0:
                    p.normal_(0, 0.05)
        mb = x070.Block(args, layer_id)
        missing = mb.load_state_dict(rb.state_dict(), strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        x = torch.randn(B, T, C)
        mask = torch.ones(B, T, 1, dtype=torch.bool)
        mask[0, :3] = False                                # left padding of sample 0
        v_first = torch.randn(B, T, C)
        st_r = [torch.randn(B, C) * 0.1, torch.randn(B, 2, 64, 64) * 0.1, torch.randn(B, C) * 0.1]
        st_m = [t.clone() for t in st_r]
        out_r = rb.forward_batch(x, mask, v_first.clone(), st_r[0], st_r[1], st_r[2])
        out_m = mb.forward_batch(x, mask, v_first.clone(), st_m[0], st_m[1], st_m[2])
        names = ("x", "v_first", "tx_prev", "state", "cx_prev")
        for n, a, b in zip(names, out_r, out_m):
            assert a.shape == b.shape and a.dtype == b.dtype, n
            err = float((a.float() - b.float()).norm() / a.float().norm().clamp(min=1e-9))
            assert err < 1e-4, (layer_id, n, err)          # fp32 modules; with kk masked too the states are off by 4e-2
        assert out_m[3] is st_m[1]                         # recurrent state advanced in place
        assert torch.equal(st_m[1], out_m[3]) and float((st_r[1] - st_m[1]).norm() / st_r[1].norm()) < 1e-4
        # second call continues from the returned states (decode step, T = 1)
        x1 = torch.randn(B, 1, C)
        m1 = torch.ones(B, 1, 1, dtype=torch.bool)
        r2 = rb.forward_batch(x1, m1, out_r[1][:, -1:], out_r[2], out_r[3], out_r[4])
        m2 = mb.forward_batch(x1, m1, out_m[1][:, -1:], out_m[2], out_m[3], out_m[4])
        for n, a, b in zip(names, r2, m2):
            err = float((a.float() - b.float()).norm() / a.float().norm().clamp(min=1e-9))
            assert err < 1e-4, (layer_id, "step", n, err)


def test_exact_mode_chain_is_operation_for_operation_the_reference_chain(monkeypatch):
    """bf16 on both sides, bit-for-bit: with core.EXACT the ATen chain around the op performs the reference's operations
    in the reference's order (same roundings), which is what makes greedy token ids of generate(exact=True) identical to
    the reference decode loop once the op itself is bit-identical (csrc/wkv7_step_exact.cu, checked on the GPU)."""
    from rwkvtts_b200 import core, x070
    ref = _reference_classes()

    def op_bf16(state, r, w, k, v, a, b):
        from oracle.wkv7_oracle import wkv7_state_forward
        y, sT = wkv7_state_forward(state.double(), *(t.double() for t in (r, w, k, v, a, b)))
        state.copy_(sT.to(state.dtype))
        return y.to(torch.bfloat16)

    ref["RWKV_Tmix_x070"].forward_batch.__globals__["RWKV7_BATCH_OP"] = op_bf16

    def wkv(r, w, k, v, a, b, state, need_state, inplace_state=False):
        st = state if inplace_state else state.clone()
        return op_bf16(st, r, w, k, v, a, b), st

    monkeypatch.setattr(core, "_wkv", wkv)
    monkeypatch.setattr(core, "EXACT", True)
    monkeypatch.setattr(core, "FUSED", False)
    args = Namespace(n_embd=128, n_layer=3, head_size_a=64, head_size_divisor=8, dim_att=128, dim_ffn=512, dropout=0.0,
                     need_init_tmix=True, need_init_cmix=True)
    torch.manual_seed(1)
    B, T, C = 3, 5, 128
    for layer_id in (0, 1):
        rb = ref["Block"](args, layer_id)
        with torch.no_grad():
            for p in rb.parameters():
                if float(p.abs().sum()) == 0:
                    p.normal_(0, 0.05)
        rb = rb.to(torch.bfloat16)
        mb = x070.Block(args, layer_id).to(torch.bfloat16)
        mb.load_state_dict(rb.state_dict(), strict=True)
        x = torch.randn(B, T, C).bfloat16()
        mask = torch.ones(B, T, 1, dtype=torch.bfloat16)
        v_first = torch.randn(B, T, C).bfloat16()
        st_r = [(torch.randn(B, C) * 0.1).bfloat16(), torch.randn(B, 2, 64, 64) * 0.1, (torch.randn(B, C) * 0.1).bfloat16()]
        st_m = [t.clone() for t in st_r]
        out_r = rb.forward_batch(x, mask, v_first.clone(), st_r[0], st_r[1], st_r[2])
        out_m = mb.forward_batch(x, mask, v_first.clone(), st_m[0], st_m[1], st_m[2])
        for n, a, b in zip(("x", "v_first", "tx_prev", "state", "cx_prev"), out_r, out_m):
            assert torch.equal(a, b), (layer_id, n, float((a.float() - b.float()).abs().max()))
        x1 = torch.randn(B, 1, C).bfloat16()
        m1 = torch.ones(B, 1, 1, dtype=torch.bfloat16)
        r2 = rb.forward_batch(x1, m1, out_r[1][:, -1:], out_r[2], out_r[3], out_r[4])
        m2 = mb.forward_batch(x1, m1, out_m[1][:, -1:], out_m[2], out_m[3], out_m[4])
        for n, a, b in zip(("x", "v_first", "tx_prev", "state", "cx_prev"), r2, m2):
            assert torch.equal(a, b), (layer_id, "step", n)
