"""GPU parity of the module stacks around the WKV op against golden
vectors produced by the REFERENCE's own Python (tests/golden/make_golden.py) and against the oracle.

Tolerance: the goldens are fp32 end to end; here the recurrence takes bf16 inputs and returns bf16
(as the reference's CUDA op does), so outputs are compared in relative L2 with a bf16-sized bar."""
import os
from argparse import Namespace

import pytest
import torch

from oracle import wkv7_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TMIX_ORDER = ["x_r", "x_w", "x_k", "x_v", "x_a", "x_g", "w0", "w1", "w2", "a0", "a1", "a2", "v0", "v1", "v2",
              "g1", "g2", "k_k", "k_a", "r_k", "R_", "K_", "V_", "O_", "ln_w", "ln_b"]
BAR = 1.5e-2


@pytest.fixture(scope="module")
def X():
    from rwkvtts_b200 import x070
    assert torch.cuda.is_available()
    return x070


def _cuda(t):
    return t.cuda() if torch.is_tensor(t) else t


@pytest.mark.parametrize("layer_id", [0, 1])
def test_block_forward_vs_reference_golden(X, layer_id):
    g = torch.load(f"{GOLD}/block_L{layer_id}.pt")
    args = Namespace(**g["args"])
    blk = X.Block(args, layer_id)
    blk.load_state_dict(g["state_dict"])          # same parameter names and shapes as the reference
    blk = blk.cuda()
    with torch.no_grad():
        y, vf = blk(g["x"].cuda(), g["mask"].cuda(), g["v_first_in"].cuda())
    e = O.rel_l2(y.float().cpu(), g["y"])
    print(f"block L{layer_id}: rel-L2 {e:.2e}")
    assert e < BAR
    assert O.rel_l2(vf.float().cpu(), g["v_first_out"]) < 1e-4


@pytest.mark.parametrize("layer_id", [0, 1])
def test_block_backward_matches_fp64_oracle_gradients(X, layer_id):
    """d(sum(y*dy))/dx through Block: autograd over our CUDA backward vs autograd over the fp64 oracle."""
    from oracle import rwkv7_model_oracle as MO
    g = torch.load(f"{GOLD}/block_L{layer_id}.pt")
    args = Namespace(**g["args"])
    blk = X.Block(args, layer_id)
    blk.load_state_dict(g["state_dict"])
    blk = blk.cuda()
    x = g["x"].cuda().requires_grad_(True)
    y, _ = blk(x, g["mask"].cuda(), g["v_first_in"].cuda())
    dy = torch.randn(y.shape, generator=torch.Generator().manual_seed(3)).cuda()
    (y * dy).sum().backward()
    # finite-difference check of the same scalar along a random direction, in the fp64 oracle
    d = torch.randn(g["x"].shape, generator=torch.Generator().manual_seed(4), dtype=torch.float64)
    eps = 1e-5
    H, N = args.n_embd // args.head_size_a, args.head_size_a
    f = lambda xx: (MO.block_train(g["state_dict"], layer_id, H, N, xx, g["mask"], g["v_first_in"],
                                   args.head_size_divisor)[0] * dy.double().cpu()).sum()
    fd = (f(g["x"].double() + eps * d) - f(g["x"].double() - eps * d)) / (2 * eps)
    an = (x.grad.double().cpu() * d).sum()
    assert abs(fd - an) < 3e-2 * max(1.0, abs(fd)), (float(fd), float(an))


@pytest.mark.parametrize("layer_id", [0, 1])
def test_tmix_one_and_seq_vs_reference_golden(X, layer_id):
    """a6: T decode steps of RWKV_x070_TMix_one == the reference's own T steps (outputs, x_prev, state)."""
    g = torch.load(f"{GOLD}/tmix_one_L{layer_id}.pt")
    H, N = g["H"], g["N"]
    w = [g["weights"][n].cuda() for n in TMIX_ORDER]
    T = g["x"].shape[0]
    # (i) step by step
    x_prev, state = g["x_prev0"].cuda(), g["state0"].cuda().clone()
    outs, vfs = [], []
    for t in range(T):
        o, x_prev, state, vf = X.RWKV_x070_TMix_one(layer_id, H, N, g["x"][t].cuda(), x_prev,
                                                    g["v_first_in"][t].cuda(), state, *w)
        outs.append(o)
        vfs.append(vf)
    out = torch.stack(outs).float().cpu()
    e = O.rel_l2(out, g["out"])
    print(f"tmix_one L{layer_id}: rel-L2 {e:.2e}")
    assert e < BAR
    assert O.rel_l2(state.cpu(), g["state_T"]) < BAR
    assert torch.equal(x_prev.cpu(), g["x_prev_T"])
    assert O.rel_l2(torch.stack(vfs).float().cpu(), g["v_first_out"]) < 1e-4
    # (ii) all T tokens in one call
    state2 = g["state0"].cuda().clone()
    o2, xl, state2, vf2 = X.RWKV_x070_TMix_seq(layer_id, H, N, g["x"].cuda(), g["x_prev0"].cuda(),
                                               g["v_first_in"].cuda(), state2, *w)
    assert O.rel_l2(o2.float().cpu(), g["out"]) < BAR
    assert O.rel_l2(state2.cpu(), g["state_T"]) < BAR


def test_cmix_one_vs_reference_golden(X):
    g = torch.load(f"{GOLD}/cmix_one.pt")
    x_prev = g["x_prev0"].cuda()
    outs = []
    for t in range(g["x"].shape[0]):
        o, x_prev = X.RWKV_x070_CMix_one(g["x"][t].cuda(), x_prev, g["x_k"].cuda(), g["K_"].cuda(), g["V_"].cuda())
        outs.append(o)
    assert O.rel_l2(torch.stack(outs).cpu(), g["out"]) < 1e-3        # TF32-free fp32 GEMMs
    o2, xl = X.RWKV_x070_CMix_seq(g["x"].cuda(), g["x_prev0"].cuda(), g["x_k"].cuda(), g["K_"].cuda(), g["V_"].cuda())
    assert O.rel_l2(o2.cpu(), g["out"]) < 1e-3


# -------------------------------------------------------------------------------------------------
# rwkvfla seam
# -------------------------------------------------------------------------------------------------
def _tiny_lm(layers=2, hidden=128, vocab=97, seed=0):
    from rwkvfla.models.rwkv7 import RWKV7Config, RWKV7ForCausalLM
    torch.manual_seed(seed)
    cfg = RWKV7Config(hidden_size=hidden, num_hidden_layers=layers, vocab_size=vocab, decay_low_rank_dim=32,
                      a_low_rank_dim=32, v_low_rank_dim=32, gate_low_rank_dim=32, fuse_cross_entropy=True)
    m = RWKV7ForCausalLM(cfg)
    with torch.no_grad():      # zero-initialised projections would hide the recurrence: make them live
        for n, p in m.named_parameters():
            if p.abs().sum() == 0 or "embeddings" in n:
                p.copy_(torch.randn_like(p) * 0.05)
    return m.cuda().to(torch.bfloat16)


def test_fla_attention_equals_x070_tmix_under_the_reference_name_map(X):
    """utils/convert_rwkv.py:17-41 maps rwkvfla names to BlinkDL names; both stacks must then agree."""
    from rwkvfla.layers.rwkv7 import RWKV7Attention
    torch.manual_seed(1)
    C, L = 128, 3
    att = RWKV7Attention(hidden_size=C, layer_idx=1, num_hidden_layers=L, decay_low_rank_dim=32,
                         a_low_rank_dim=32, v_low_rank_dim=32, gate_low_rank_dim=128)
    with torch.no_grad():
        for p in att.parameters():
            if p.abs().sum() == 0:
                p.copy_(torch.randn_like(p) * 0.05)
    args = Namespace(n_layer=L, n_embd=C, head_size_a=64, head_size_divisor=8, dropout=0.0)
    tm = X.RWKV_Tmix_x070(args, 1)
    sd = {}
    f = dict(att.state_dict())
    for n in ("x_r", "x_w", "x_k", "x_v", "x_a", "x_g"):
        sd[n] = f[n]
    sd["k_k"], sd["k_a"], sd["r_k"] = f["k_k"].view(1, 1, C), f["k_a"].view(1, 1, C), f["r_k"]
    for a_, b_ in (("receptance", "r_proj"), ("key", "k_proj"), ("value", "v_proj"), ("output", "o_proj")):
        sd[a_ + ".weight"] = f[b_ + ".weight"]
    for c in "wavg":
        sd[c + "1"] = f[f"{c}_lora.lora.0.weight"].t()
        sd[c + "2"] = f[f"{c}_lora.lora.2.weight"].t()
        if c != "g":
            sd[c + "0"] = f[f"{c}_lora.lora.2.bias"].view(1, 1, C)
    sd["ln_x.weight"], sd["ln_x.bias"] = f["g_norm.weight"], f["g_norm.bias"]
    tm.load_state_dict({k: v.reshape(tm.state_dict()[k].shape) for k, v in sd.items()})
    att, tm = att.cuda(), tm.cuda()
    x = torch.randn(2, 32, C, device="cuda")
    vf = torch.randn(2, 32, C, device="cuda")
    with torch.no_grad():
        o1, _, _, _ = att(x, v_first=vf)
        o2, _ = tm(x, None, vf)
    assert O.rel_l2(o1.cpu(), o2.cpu()) < 1e-3


def test_causal_lm_prefill_plus_steps_equals_full_forward():
    m = _tiny_lm()
    m.eval()
    ids = torch.randint(0, 97, (3, 37), device="cuda")
    from rwkvfla.models.utils import Cache
    with torch.no_grad():
        full = m(input_ids=ids).logits.float()
        cache = Cache()
        out = m(input_ids=ids[:, :32], past_key_values=cache, use_cache=True)          # chunked prefill, T = 32
        steps = [out.logits.float()]
        for t in range(32, 37):                                                       # decode steps
            out = m(input_ids=ids[:, t:t + 1], past_key_values=out.past_key_values, use_cache=True)
            steps.append(out.logits.float())
    inc = torch.cat(steps, dim=1)
    assert O.rel_l2(inc.cpu(), full.cpu()) < 2e-2
    assert out.past_key_values.seen_tokens == 37
    st = out.past_key_values[0]
    assert st["recurrent_state"].shape == (3, 2, 64, 64) and st["recurrent_state"].dtype == torch.float32
    assert st["conv_state"].shape == (3, 128) and st["ffn_state"].shape == (3, 128)


def test_left_padding_matches_unpadded():
    m = _tiny_lm(seed=2)
    m.eval()
    ids = torch.randint(0, 97, (1, 20), device="cuda")
    pad = torch.zeros(1, 12, dtype=torch.long, device="cuda")
    padded = torch.cat([pad, ids], dim=1)
    mask = torch.cat([torch.zeros(1, 12), torch.ones(1, 20)], dim=1).cuda()
    with torch.no_grad():
        a = m(input_ids=ids).logits.float()
        b = m(input_ids=padded, attention_mask=mask).logits.float()[:, 12:]
    assert O.rel_l2(b.cpu(), a.cpu()) < 2e-2


def test_generate_greedy_is_deterministic_and_matches_manual_argmax_loop():
    m = _tiny_lm(seed=3)
    m.eval()
    ids = torch.randint(0, 97, (4, 16), device="cuda")
    g1 = m.generate(input_ids=ids, max_new_tokens=12, do_sample=False, eos_token_id=None)
    g2 = m.generate(input_ids=ids, max_new_tokens=12, do_sample=False, eos_token_id=None)
    assert torch.equal(g1, g2) and g1.shape == (4, 28) and torch.equal(g1[:, :16], ids)
    emb = m.get_input_embeddings()(ids)
    g3 = m.generate(inputs_embeds=emb, max_new_tokens=12, do_sample=False, eos_token_id=None)
    assert g3.shape == (4, 12) and torch.equal(g3, g1[:, 16:])
    s = m.generate(input_ids=ids, max_new_tokens=8, do_sample=True, top_k=5, top_p=0.9, temperature=0.8,
                   generator=torch.Generator(device="cuda").manual_seed(0), eos_token_id=None)
    assert s.shape == (4, 24)


def test_generate_cuda_graph_step_is_bit_exact_with_the_eager_step():
    """greedy token ids must not depend on whether the per-token step runs eagerly or as a CUDA graph (BASELINE: bit-exact
    argmax ids under greedy decode), including continuing from the returned cache"""
    m = _tiny_lm(seed=5)
    m.eval()
    ids = torch.randint(0, 97, (3, 16), device="cuda")
    a = m.generate(input_ids=ids, max_new_tokens=40, do_sample=False, eos_token_id=None, use_cuda_graph=False)
    b = m.generate(input_ids=ids, max_new_tokens=40, do_sample=False, eos_token_id=None, use_cuda_graph=True,
                   use_megakernel=False, return_dict_in_generate=True)
    assert torch.equal(a, b["sequences"])
    assert b["past_key_values"].seen_tokens == 16 + 39
    # eos handling: rows that hit eos are padded, the others keep decoding; polling interval does not change the result
    eos = int(a[0, 20])
    c = m.generate(input_ids=ids, max_new_tokens=40, do_sample=False, eos_token_id=eos, pad_token_id=0, use_cuda_graph=False)
    d = m.generate(input_ids=ids, max_new_tokens=40, do_sample=False, eos_token_id=eos, pad_token_id=0, use_cuda_graph=True,
                   use_megakernel=False, eos_check_interval=8)
    n = min(c.shape[1], d.shape[1])
    assert torch.equal(c[:, :n], d[:, :n]) and bool((d[:, n:] == 0).all() if d.shape[1] > n else True)


def test_training_step_backward_gives_finite_gradients_for_every_parameter():
    m = _tiny_lm(seed=4)
    m.train()
    ids = torch.randint(0, 97, (2, 48), device="cuda")
    labels = ids.clone()
    labels[:, :10] = -100
    out = m(input_ids=ids, labels=labels)
    assert torch.isfinite(out.loss)
    out.loss.backward()
    for n, p in m.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad.float()).all(), n
    # fused linear + CE path gives the same loss
    m.config.fuse_linear_cross_entropy = True
    out2 = m(input_ids=ids, labels=labels)
    assert abs(float(out2.loss) - float(out.loss)) < 2e-2 * abs(float(out.loss))


# ---------------------------------------------------------------------------------------------------------------
# greedy-id parity against the reference decode step (north_star: bit-exact argmax ids; round-1 VERDICT missing #3)
# ---------------------------------------------------------------------------------------------------------------
def _ref_step_available():
    from oracle import c_oracle as CO
    return CO.ref_available()


@pytest.mark.parametrize("B,T,H", [(32, 1, 16), (3, 163, 4), (1, 16, 2), (5, 64, 1)])
def test_exact_step_kernel_is_bit_identical_to_the_reference_kernel(B, T, H):
    """rwkvtts_set_step_mode(1): y and the recurrent state equal the UNMODIFIED reference kernel's
    (rwkv7_state_fwd_fp16.cu, compiled in oracle/_ref) bit for bit, decode step (T = 1) and prefill (any T)."""
    if not _ref_step_available():
        pytest.skip("oracle/_ref not built")
    import rwkvtts_b200 as R
    from oracle import c_oracle as CO
    from oracle import wkv7_oracle as O
    x = O.make_inputs(B, T, H, seed=B * 100 + T)
    flat = {n: t.cuda().view(B, T, H * 64) for n, t in x.items()}
    s0 = (torch.randn(B, H, 64, 64, generator=torch.Generator().manual_seed(1)) * 0.3).cuda()
    st_ref, st_our = s0.clone(), s0.clone()
    args = [flat[n] for n in "qwkvab"]
    y_ref = CO.ref_state_forward(st_ref, *args)
    L = R._lib.lib()
    assert L.rwkvtts_set_step_mode(1) == 0
    try:
        y_our = R.RWKV7_BATCH_OP(st_our, *args)
    finally:
        L.rwkvtts_set_step_mode(0)
    torch.cuda.synchronize()
    assert torch.equal(y_our, y_ref)
    assert torch.equal(st_our, st_ref)
    # the fast kernels agree to rounding, not to the bit (which is why the exact mode exists)
    st_fast = s0.clone()
    y_fast = R.RWKV7_BATCH_OP(st_fast, *args)
    assert float((y_fast.float() - y_ref.float()).norm() / y_ref.float().norm()) < 4e-3


def test_generate_exact_mode_reproduces_the_reference_greedy_loop():
    """Small model, 8 prompts x 192 greedy steps: generate(exact=True) against the reference decode loop rebuilt from the
    reference kernel + the ATen chain (scripts/decode_parity.py) -- identical ids at every step; the default fast path is
    compared the same way and may only differ where the reference's own top-2 logits are within bf16 resolution."""
    if not _ref_step_available():
        pytest.skip("oracle/_ref not built")
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import decode_parity
    from rwkvfla.models.rwkv7 import RWKV7Config, RWKV7ForCausalLM
    torch.manual_seed(3)
    cfg = RWKV7Config(hidden_size=256, num_hidden_layers=3, head_dim=64, vocab_size=1025, decay_low_rank_dim=32,
                      a_low_rank_dim=32, v_low_rank_dim=16, gate_low_rank_dim=64)
    m = RWKV7ForCausalLM(cfg)
    with torch.no_grad():
        for _, p in m.named_parameters():
            if p.abs().sum() == 0:
                p.copy_(torch.randn_like(p) * 0.05)
    m = m.cuda().to(torch.bfloat16).eval()
    ids = torch.randint(0, 1024, (8, 37), device="cuda")
    NEW = 192
    for graph in (False, True):
        seq = m.generate(input_ids=ids, max_new_tokens=NEW, do_sample=False, eos_token_id=None, exact=True, use_cuda_graph=graph)
        res = decode_parity.compare(m, ids, seq[:, 37:], NEW)
        assert res["identical"], res
    fast = m.generate(input_ids=ids, max_new_tokens=NEW, do_sample=False, eos_token_id=None)
    res = decode_parity.compare(m, ids, fast[:, 37:], NEW)
    print("fast path vs reference loop:", res)
    assert res["mismatches"] <= 0.1 * res["ids_compared"]
    assert res["max_reference_logit_gap_at_mismatch"] < 0.25        # only near-ties of the reference's own logits flip
