"""Caller side of the path on the GPU: the embedding gather kernel behind the batch
builders and the fused linear + cross-entropy head, against the same lines in plain PyTorch."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


class _Model(torch.nn.Module):
    def __init__(self, D=128):
        super().__init__()
        self.tts_tag_embedder = torch.nn.Embedding(3, D)
        self.text_embedder = torch.nn.Embedding(500, D)
        self.global_embedder = torch.nn.Embedding(64, D)
        self.model = torch.nn.Module()
        self.model.embeddings = torch.nn.Embedding(131, D)

    @property
    def device(self):
        return self.text_embedder.weight.device


def _spark_batch(B=5, seed=0):
    g = torch.Generator().manual_seed(seed)
    lt, lg, ls = 9, 6, 40
    b = {"input_ids": torch.randint(0, 500, (B, lt), generator=g), "global_tokens_ids": torch.randint(0, 64, (B, lg), generator=g),
         "semantic_tokens_ids": torch.randint(0, 130, (B, ls), generator=g)}
    def left_mask(L, lens):
        return (torch.arange(L)[None, :] >= (L - torch.tensor(lens))[:, None]).long()
    b["attention_mask_input_ids"] = left_mask(lt, [9, 3, 1, 7, 9][:B])
    b["global_tokens_attention_mask"] = left_mask(lg, [6, 6, 2, 1, 5][:B])
    b["semantic_tokens_attention_mask"] = left_mask(ls, [40, 11, 1, 33, 25][:B])
    return {k: v.cuda() for k, v in b.items()}


def test_gather_kernel_equals_lookup_and_scatter_in_every_builder():
    from rwkvtts_b200 import batch as BT, core
    torch.manual_seed(1)
    m = _Model().cuda().to(torch.bfloat16)
    b = _spark_batch()
    res = {}
    for fused_on in (True, False):
        core.FUSED = fused_on
        try:
            m.zero_grad(set_to_none=True)
            o1 = BT.process_single_batch(b, m, eos_token_id=130)
            o2 = BT.process_single_batch_culens(b, m, eos_token_id=130, max_cu_seqlens=4096)
            (o1["input_embs"].float().square().sum() + o2["input_embs"].float().mul(0.5).sum()).backward()
            res[fused_on] = (o1, o2, {n: p.grad.clone() for n, p in m.named_parameters()})
        finally:
            core.FUSED = True
    for a, c in zip(res[True][:2], res[False][:2]):
        for k in a:
            assert torch.equal(a[k], c[k]), k                     # pure copies: bit-identical
    for n in res[True][2]:
        assert torch.allclose(res[True][2][n].float(), res[False][2][n].float(), rtol=2e-2, atol=1e-3), n


def test_gather_kernel_launch_count_is_one():
    from rwkvtts_b200 import _lib, batch as BT
    m = _Model().cuda().to(torch.bfloat16)
    b = _spark_batch()
    n0 = _lib.lib().rwkvtts_kernel_launches()
    with torch.no_grad():
        BT.process_single_batch(b, m, eos_token_id=130)
    assert _lib.lib().rwkvtts_kernel_launches() - n0 == 1


@pytest.mark.parametrize("V,N,D,eps,red", [(8193, 700, 256, 0.0, "mean"), (131, 64, 64, 0.0, "sum"), (6562, 300, 128, 0.1, "mean"),
                                            (66690, 130, 128, 0.0, "mean"), (16384, 65, 64, 0.0, "mean")])
def test_linear_cross_entropy_matches_torch(V, N, D, eps, red):
    from rwkvtts_b200 import fused
    g = torch.Generator(device="cuda").manual_seed(V)
    h = (torch.randn(N, D, device="cuda", generator=g) * 0.7).bfloat16().requires_grad_(True)
    W = (torch.randn(V, D, device="cuda", generator=g) * 0.2).bfloat16().requires_grad_(True)
    y = torch.randint(0, V, (N,), device="cuda", generator=g)
    y[::7] = -100
    loss = fused.linear_cross_entropy(h, y, W, ignore_index=-100, label_smoothing=eps, reduction=red, chunk_rows=256)
    (loss * 1.5).backward()
    hr, Wr = h.detach().float().requires_grad_(True), W.detach().float().requires_grad_(True)
    ref = F.cross_entropy(hr @ Wr.t(), y, ignore_index=-100, label_smoothing=eps, reduction=red)
    (ref * 1.5).backward()
    # the logits are rounded to bf16 by the GEMM (the reference's lm_head output is bf16 as well)
    assert abs(float(loss) - float(ref)) < 3e-3 * abs(float(ref)), (float(loss), float(ref))
    rel = lambda a, b: float((a.float() - b).norm() / b.norm())
    assert rel(h.grad, hr.grad) < 1.5e-2, rel(h.grad, hr.grad)
    assert rel(W.grad, Wr.grad) < 1.5e-2, rel(W.grad, Wr.grad)
    assert float(h.grad[0].abs().sum()) == 0.0            # ignored row


def test_fused_linear_ce_module_uses_the_kernel_and_equals_the_chunked_path():
    from rwkvfla.modules import FusedLinearCrossEntropyLoss
    from rwkvtts_b200 import _lib, core
    g = torch.Generator(device="cuda").manual_seed(5)
    h = torch.randn(3, 50, 128, device="cuda", generator=g).bfloat16()
    W = (torch.randn(8193, 128, device="cuda", generator=g) * 0.1).bfloat16()
    y = torch.randint(0, 8193, (3, 50), device="cuda", generator=g)
    y[:, :5] = -100
    crit = FusedLinearCrossEntropyLoss()
    n0 = _lib.lib().rwkvtts_kernel_launches()
    a = crit(h, y, W)
    assert _lib.lib().rwkvtts_kernel_launches() > n0
    core.FUSED = False
    try:
        b = crit(h, y, W)
    finally:
        core.FUSED = True
    assert abs(float(a) - float(b)) < 3e-3 * abs(float(b))
