"""The one-kernel decode step (csrc/decode_step.cu, rwkvtts_b200/decode.py) against the
decode path it replaces -- the per-layer fused kernels + cuBLAS projections of the eager / CUDA-graph step -- and against
an fp32 run of the same model through the plain ATen chain.

Tolerance: both paths round to bf16 at the same points and differ in the summation order of the projections, so logits
agree to bf16 resolution compounded over the layers: relative L2 <= 2e-2 against each other, and the one-kernel step must
be as close to the fp32 chain as the path it replaces (within 1.5x).  Device-side greedy sampling is compared id for id
against the host loop over the SAME kernel's logits (deterministic, so identical)."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


def _lm(hidden, layers, vocab, ranks, seed=0, ratio=4, norm_bias=True):
    from rwkvfla.models.rwkv7 import RWKV7Config, RWKV7ForCausalLM
    torch.manual_seed(seed)
    cfg = RWKV7Config(hidden_size=hidden, num_hidden_layers=layers, vocab_size=vocab, decay_low_rank_dim=ranks[0],
                      a_low_rank_dim=ranks[1], v_low_rank_dim=ranks[2], gate_low_rank_dim=ranks[3], hidden_ratio=ratio,
                      norm_bias=norm_bias)
    m = RWKV7ForCausalLM(cfg)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.abs().sum() == 0 or "embeddings" in n:
                p.copy_(torch.randn_like(p) * 0.05)
    return m.cuda().to(torch.bfloat16).eval()


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def _prefill(m, ids):
    from rwkvfla.models.utils import Cache
    with torch.no_grad():
        out = m(input_ids=ids, past_key_values=Cache(), use_cache=True, logits_to_keep=1)
    return out.past_key_values, out.logits[:, -1].float()


def _clone_cache(c):
    n = copy.copy(c)
    n.states = [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in st.items()} for st in c.states]
    return n


@pytest.mark.parametrize("hidden,layers,vocab,ranks,B,ratio,bias", [
    (128, 3, 97, (32, 32, 32, 64), 5, 4, True),        # two k blocks per row, odd vocabulary, B not a multiple of 8
    (768, 2, 1000, (64, 64, 32, 128), 32, 4, True),    # the 0.1B width (12 of 16 warps hold a K slice), full batch
    (1024, 2, 8193, (64, 64, 32, 128), 1, 4, False),   # the 0.4B width and head, one row, norms without bias
    (256, 2, 130, (32, 64, 96, 32), 7, 2, True),       # uneven ranks, channel-mix width 2 C
])
def test_one_kernel_step_matches_the_fused_eager_step(hidden, layers, vocab, ranks, B, ratio, bias):
    from rwkvtts_b200 import core
    from rwkvtts_b200.decode import MegaDecodeStep
    m = _lm(hidden, layers, vocab, ranks, seed=hidden + B, ratio=ratio, norm_bias=bias)
    ids = torch.randint(0, vocab, (B, 19), device="cuda")
    cache, _ = _prefill(m, ids)
    c_mega, c_eager, c_f32 = _clone_cache(cache), _clone_cache(cache), _clone_cache(cache)
    mega = MegaDecodeStep(m, c_mega, B, torch.device("cuda"))
    m32 = copy.deepcopy(m).float()
    for st in c_f32.states:
        for k in ("conv_state", "ffn_state"):
            st[k] = st[k].float()
    worst, worst_ref = 0.0, 0.0
    g = torch.Generator(device="cuda").manual_seed(1)
    for step in range(6):
        tok = torch.randint(0, vocab, (B,), device="cuda", generator=g)
        lm = mega(tok).clone()
        with torch.no_grad():
            le = m(input_ids=tok[:, None], past_key_values=c_eager, use_cache=True, logits_to_keep=1).logits[:, -1].float()
            prev = core.FUSED
            core.FUSED = False
            try:
                lf = m32(input_ids=tok[:, None], past_key_values=c_f32, use_cache=True, logits_to_keep=1).logits[:, -1].float()
            finally:
                core.FUSED = prev
        assert torch.isfinite(lm).all()
        e, em, ee = _rel(lm, le), _rel(lm, lf), _rel(le, lf)
        print(f"step {step}: one-kernel vs eager {e:.2e}; vs fp32 chain: one-kernel {em:.2e}, eager {ee:.2e}")
        worst, worst_ref = max(worst, e), max(worst_ref, em / max(ee, 1e-4))
    assert worst < 2e-2
    assert worst_ref < 1.5
    for l, (a, b) in enumerate(zip(c_mega.states, c_eager.states)):
        assert _rel(a["recurrent_state"], b["recurrent_state"]) < 2e-2, l
        assert _rel(a["conv_state"].float(), b["conv_state"].float()) < 2e-2, l
        assert _rel(a["ffn_state"].float(), b["ffn_state"].float()) < 2e-2, l
    assert c_mega.seen_tokens == c_eager.seen_tokens


def test_one_kernel_step_is_deterministic_and_survives_many_launches():
    from rwkvtts_b200.decode import MegaDecodeStep
    m = _lm(256, 3, 513, (32, 32, 32, 64), seed=11)
    ids = torch.randint(0, 513, (9, 12), device="cuda")
    cache, _ = _prefill(m, ids)
    runs = []
    for _ in range(2):
        c = _clone_cache(cache)
        mega = MegaDecodeStep(m, c, 9, torch.device("cuda"))
        tok = ids[:, -1].clone()
        acc = []
        for _ in range(300):                                    # 300 launches x 24 grid barriers each
            lg = mega(tok)
            tok = lg.argmax(dim=-1)
            acc.append(tok)
        runs.append((torch.stack(acc), lg.clone(), c.states[-1]["recurrent_state"].clone()))
        mega.close()
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1]) and torch.equal(runs[0][2], runs[1][2])


def _host_loop(m, ids, new, eos, pad, min_new):
    """generate()'s greedy loop on the host, over the one-kernel step's logits."""
    from rwkvtts_b200.decode import MegaDecodeStep
    cache, logits = _prefill(m, ids)
    B = ids.shape[0]
    mega = MegaDecodeStep(m, cache, B, ids.device)
    eos_t = torch.tensor(eos, device=ids.device, dtype=torch.long) if eos else None
    done = torch.zeros(B, dtype=torch.bool, device=ids.device)
    toks = []
    for step in range(new):
        if eos_t is not None and step < min_new:
            logits[:, eos_t] = float("-inf")
        nxt = logits.argmax(dim=-1)
        nxt = torch.where(done, torch.full_like(nxt, pad), nxt)
        toks.append(nxt)
        if eos_t is not None:
            done = done | torch.isin(nxt, eos_t)
            if bool(done.all()):
                break
        if step + 1 < new:
            logits = mega(nxt).clone()
    return torch.stack(toks, dim=1)


def test_device_greedy_sampling_equals_the_host_loop():
    m = _lm(128, 2, 97, (32, 32, 32, 32), seed=5)
    ids = torch.randint(0, 97, (6, 16), device="cuda")
    free = m.generate(input_ids=ids, max_new_tokens=150, do_sample=False, eos_token_id=None, use_megakernel=True)
    assert free.shape == (6, 166) and torch.equal(free[:, :16], ids)
    assert torch.equal(free[:, 16:], _host_loop(m, ids, 150, [], 0, 0))
    # EOS ids that do occur: rows finish at different times, are padded, and the sequence is cut where the host loop stops
    gen = free[:, 16:]
    eos = [int(gen[0, 10]), int(gen[3, 40])]
    for min_new in (0, 30):
        want = _host_loop(m, ids, 150, eos, 96, min_new)
        got = m.generate(input_ids=ids, max_new_tokens=150, do_sample=False, eos_token_id=eos, pad_token_id=96,
                         min_new_tokens=min_new, use_megakernel=True)[:, 16:]
        assert got.shape == want.shape and torch.equal(got, want), (min_new, got.shape, want.shape)
    # sampling path: logits come from the one-kernel step, sampling stays on the host
    s = m.generate(input_ids=ids, max_new_tokens=20, do_sample=True, top_k=5, temperature=0.9, eos_token_id=None,
                   generator=torch.Generator(device="cuda").manual_seed(0), use_megakernel=True)
    assert s.shape == (6, 36)


def test_one_kernel_generate_agrees_with_the_graph_step_where_logits_are_not_tied():
    """Greedy ids of the two fast paths on a small model: they differ only by summation order, so the first step at which
    they pick different ids must be a near-tie of the top-2 logits."""
    m = _lm(256, 3, 257, (32, 32, 32, 64), seed=2)
    ids = torch.randint(0, 257, (8, 24), device="cuda")
    a = m.generate(input_ids=ids, max_new_tokens=64, do_sample=False, eos_token_id=None, use_megakernel=True)[:, 24:]
    b = m.generate(input_ids=ids, max_new_tokens=64, do_sample=False, eos_token_id=None, use_megakernel=False)[:, 24:]
    same = (a == b).long().cumprod(dim=1).sum(dim=1)            # tokens until the first disagreement, per row
    print("identical prefix lengths:", same.tolist())
    assert int(same.min()) >= 4


def test_unsupported_models_fall_back_and_say_why():
    from rwkvtts_b200.decode import MegaDecodeStep, unsupported_reason
    m = _lm(128, 2, 97, (32, 32, 16, 32), seed=1)                # a rank of 16: outside the kernel's range
    assert "LoRA ranks" in unsupported_reason(m, 4)
    assert unsupported_reason(m, 33) is not None
    ids = torch.randint(0, 97, (2, 8), device="cuda")
    out = m.generate(input_ids=ids, max_new_tokens=12, do_sample=False, eos_token_id=None)       # auto: graph step
    assert out.shape == (2, 20)
    with pytest.raises(ValueError):
        m.generate(input_ids=ids, max_new_tokens=12, do_sample=False, eos_token_id=None, use_megakernel=True)
