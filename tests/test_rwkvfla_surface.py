"""Import surface and host logic of the rwkvfla / deepspeed stand-ins (no GPU needed): every symbol the
reference imports exists, parameter names match rwkv-fla's, Cache / loss modules behave."""
import pytest
import torch


def test_symbols_the_reference_imports():
    from rwkvfla.models.rwkv7.modeling_rwkv7 import (Cache, FusedCrossEntropyLoss, FusedLinearCrossEntropyLoss,  # noqa
                                                      RWKV7Block, RWKV7ForCausalLM, RWKV7Model, RWKV7PreTrainedModel)
    from rwkvfla.models.rwkv7.configuration_rwkv7 import RWKV7Config  # noqa
    from rwkvfla.models.rwkv7 import RWKV7Model as M2  # noqa
    from rwkvfla.models.utils import Cache as C2  # noqa
    from rwkvfla.layers.rwkv7 import RWKV7Attention  # noqa
    from rwkvfla.layers.rwkv6 import LoRA  # noqa
    from rwkvfla.modules import LayerNorm  # noqa
    from rwkvfla.modules.l2warp import l2_warp  # noqa
    import deepspeed  # noqa
    from deepspeed.ops.adam import DeepSpeedCPUAdam, FusedAdam  # noqa
    assert callable(deepspeed.initialize) and callable(deepspeed.init_distributed)
    assert callable(deepspeed.checkpointing.checkpoint)


def test_parameter_names_match_rwkv_fla_and_convert_map():
    from rwkvfla.models.rwkv7 import RWKV7Config, RWKV7ForCausalLM
    cfg = RWKV7Config(hidden_size=128, num_hidden_layers=2, vocab_size=50, num_heads=32)   # stale num_heads: head_dim wins
    assert cfg.num_heads == 2 and cfg.head_dim == 64
    m = RWKV7ForCausalLM(cfg)
    names = {n for n, _ in m.named_parameters()}
    for n in ["model.embeddings.weight", "model.layers.0.pre_norm.weight", "model.layers.0.attn_norm.bias",
              "model.layers.1.attn.x_r", "model.layers.1.attn.k_k", "model.layers.1.attn.k_a", "model.layers.1.attn.r_k",
              "model.layers.1.attn.r_proj.weight", "model.layers.1.attn.o_proj.weight",
              "model.layers.1.attn.w_lora.lora.0.weight", "model.layers.1.attn.w_lora.lora.2.weight",
              "model.layers.1.attn.w_lora.lora.2.bias", "model.layers.1.attn.v_lora.lora.2.bias",
              "model.layers.1.attn.a_lora.lora.2.bias", "model.layers.1.attn.g_lora.lora.2.weight",
              "model.layers.1.attn.g_norm.weight", "model.layers.1.ffn.x_k", "model.layers.1.ffn.key.weight",
              "model.layers.1.ffn.value.weight", "model.layers.1.ffn_norm.weight", "model.norm.weight", "lm_head.weight"]:
        assert n in names, n
    assert "model.layers.0.attn.v_lora.lora.0.weight" not in names       # layer 0 has no value residual
    assert "model.layers.1.attn.g_lora.lora.2.bias" not in names
    # the optimizer grouping of train_spark_rwkv7speech_jsonl.py:161-172 finds its lr_2x parameters
    assert sum("attn.w_lora.lora.2.bias" in n for n in names) == 2
    # BlinkDL init survived HF's post_init
    att = m.model.layers[1].attn
    assert float(att.o_proj.weight.abs().sum()) == 0.0 and abs(float(att.k_a[0]) - 1.02) < 1e-6


def test_cpu_forward_fails_loudly():
    from rwkvfla.models.rwkv7 import RWKV7Config, RWKV7ForCausalLM
    m = RWKV7ForCausalLM(RWKV7Config(hidden_size=64, num_hidden_layers=1, vocab_size=11))
    with pytest.raises(Exception):
        m(input_ids=torch.zeros(1, 16, dtype=torch.long))


def test_cache_protocol_and_layout_converters():
    from rwkvfla.models.utils import Cache
    c = Cache()
    assert len(c) == 0 and c.get_seq_length() == 0
    s = torch.arange(2 * 3 * 64 * 64, dtype=torch.float32).view(2, 3, 64, 64)
    c.update(recurrent_state=s, conv_state=torch.zeros(2, 192), layer_idx=0, offset=5)
    c.update(ffn_state=torch.ones(2, 192), layer_idx=0, offset=0)
    c.update(recurrent_state=s + 1, conv_state=torch.zeros(2, 192), layer_idx=1, offset=5)
    assert len(c) == 2 and c.seen_tokens == 5 and c[0]["ffn_state"].sum() == 384
    fla = c.to_fla_layout()
    assert torch.equal(fla[0]["recurrent_state"], s.transpose(-1, -2))
    back = Cache.from_fla_layout(fla, seen_tokens=5)
    assert torch.equal(back[1]["recurrent_state"], s + 1)
    sub = c.batch_select(torch.tensor([1]))
    assert sub[0]["recurrent_state"].shape[0] == 1
    assert Cache.from_legacy_cache(c) is c


def test_loss_modules_match_torch():
    from rwkvfla.modules import FusedCrossEntropyLoss, FusedLinearCrossEntropyLoss
    from rwkvfla.modules.l2warp import l2_warp
    torch.manual_seed(0)
    h = torch.randn(3, 10, 16, requires_grad=True)
    W = torch.randn(23, 16, requires_grad=True)
    y = torch.randint(0, 23, (3, 10))
    y[0, :4] = -100
    ref = torch.nn.functional.cross_entropy((h @ W.t()).view(-1, 23), y.view(-1), ignore_index=-100)
    gh, gW = torch.autograd.grad(ref, [h, W])
    a = FusedCrossEntropyLoss()((h @ W.t()).view(-1, 23), y.view(-1))
    b = FusedLinearCrossEntropyLoss(num_chunks=4)(h, y, W)
    assert torch.allclose(a, ref, atol=1e-6) and torch.allclose(b, ref, atol=1e-5)
    gh2, gW2 = torch.autograd.grad(b, [h, W])
    assert torch.allclose(gh2, gh, atol=1e-5) and torch.allclose(gW2, gW, atol=1e-5)
    logits = (h @ W.t())
    l2 = l2_warp(ref.detach().requires_grad_(True) * 1.0, logits)
    (gl,) = torch.autograd.grad(l2, [logits], allow_unused=True)
    assert gl is not None and int((gl != 0).sum()) == 30                   # one pulled logit per position


def test_layernorm_prenorm_form():
    from rwkvfla.modules import LayerNorm
    ln = LayerNorm(8, bias=True)
    x, r = torch.randn(2, 3, 8), torch.randn(2, 3, 8)
    y, res = ln(x, r, True)
    assert torch.allclose(res, x + r) and torch.allclose(y, torch.nn.functional.layer_norm(x + r, (8,), ln.weight, ln.bias))


def test_cu_seqlens_packed_input_runs_as_the_equivalent_padded_batch(monkeypatch):
    """RWKV7Model.forward(cu_seqlens=...) (what the reference's *_culens collators feed, utils/multiple_jsonl.py:78-135):
    packed [1, total, D] must give, per sequence, what the sequence gives on its own.  The blocks are replaced by a
    causal stand-in (running mean over time) so that the wiring is checked on CPU; the real blocks see an ordinary
    right-padded batch."""
    import torch
    from rwkvfla.models.rwkv7 import RWKV7Config, RWKV7Model
    from rwkvfla.models.rwkv7 import modeling_rwkv7 as M

    def fake_block(self, hidden_states, attention_mask=None, past_key_values=None, use_cache=False,
                   output_attentions=False, v_first=None, cu_seqlens=None, **kw):
        assert cu_seqlens is None and attention_mask is None
        t = torch.arange(1, hidden_states.shape[1] + 1, dtype=hidden_states.dtype).view(1, -1, 1)
        return hidden_states.cumsum(1) / t + 0.1 * (self.layer_idx + 1), None, past_key_values, v_first

    monkeypatch.setattr(M.RWKV7Block, "forward", fake_block)
    torch.manual_seed(0)
    cfg = RWKV7Config(hidden_size=64, num_hidden_layers=2, vocab_size=50, decay_low_rank_dim=32, a_low_rank_dim=32,
                      v_low_rank_dim=32, gate_low_rank_dim=32)
    m = RWKV7Model(cfg).eval()
    lens = [5, 1, 9, 3]
    cu = torch.tensor([0, 5, 6, 15, 18])
    x = torch.randn(1, 18, 64, requires_grad=True)
    out = m(inputs_embeds=x, cu_seqlens=cu, output_hidden_states=True)
    assert out.last_hidden_state.shape == (1, 18, 64) and all(h.shape == (1, 18, 64) for h in out.hidden_states)
    for a, b in zip(cu[:-1].tolist(), cu[1:].tolist()):
        single = m(inputs_embeds=x[:, a:b]).last_hidden_state
        assert torch.allclose(out.last_hidden_state[:, a:b], single, atol=1e-6)
    out.last_hidden_state.sum().backward()
    assert x.grad is not None and bool(torch.isfinite(x.grad).all())
    # helpers are inverse of each other and reject a cu_seqlens that does not partition the sequence
    padded, idx = M.unpack_varlen(x.detach(), cu)
    assert padded.shape == (4, 9, 64) and torch.equal(M.repack_varlen(padded, idx), x.detach())
    import pytest
    with pytest.raises(ValueError):
        M.unpack_varlen(x.detach(), torch.tensor([0, 5, 17]))


def test_sampling_filters_match_hf_warpers():
    """generate()'s top-k / top-p filter keeps exactly the tokens Hugging Face's warpers keep (the reference decodes
    through HF generate, spark_llm.py:54-102), for random logits, k and p."""
    import torch
    from transformers.generation.logits_process import TopKLogitsWarper, TopPLogitsWarper
    from rwkvfla.models.rwkv7.modeling_rwkv7 import _filter_logits
    g = torch.Generator().manual_seed(0)
    ids = torch.zeros(3, 1, dtype=torch.long)
    for _ in range(60):
        V = int(torch.randint(5, 300, (1,), generator=g))
        lg = torch.randn(3, V, generator=g) * 3
        k = int(torch.randint(1, V + 5, (1,), generator=g))
        p = float(torch.rand(1, generator=g)) * 0.98 + 0.01
        want = TopPLogitsWarper(top_p=p)(ids, TopKLogitsWarper(top_k=min(k, V))(ids, lg.clone()))
        got = _filter_logits(lg.clone(), k, p)
        assert torch.equal(torch.isinf(want), torch.isinf(got))
        assert torch.equal(want[~torch.isinf(want)], got[~torch.isinf(got)])
