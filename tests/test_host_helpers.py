"""Host-side helpers of the fused path that need no GPU: the per-forward mask conversion (core.mask3) and the batched
parameter-gradient conversion of the adjoints (fused._param_grads)."""
import torch

from rwkvtts_b200 import core, fused


def test_mask3_is_the_reference_conversion_and_runs_once_per_forward():
    """RWKV7Attention / RWKV7FeedForward take the last T columns of the 0/1 attention mask as [B, T, 1] in the activations'
    dtype (rwkv-fla: `attention_mask[:, -T:, None]`); every layer asks for the same tensor, so it is converted once."""
    m = torch.tensor([[0, 0, 1, 1, 1], [1, 1, 1, 1, 1]])
    a = core.mask3(m, 3, torch.bfloat16)
    assert a.shape == (2, 3, 1) and a.dtype == torch.bfloat16
    assert torch.equal(a[..., 0], m[:, -3:].to(torch.bfloat16))
    assert core.mask3(m, 3, torch.bfloat16) is a                       # same mask object, same version: the cached tensor
    assert core.mask3(m, 2, torch.bfloat16).shape == (2, 2, 1)         # another T: converted again
    b = core.mask3(m, 3, torch.float32)
    assert b.dtype == torch.float32 and b is not a
    m[0, 2] = 0                                                         # in-place edit moves the version counter
    c = core.mask3(m, 3, torch.float32)
    assert c is not b and float(c[0, 0, 0]) == 0.0
    m2 = m.clone()                                                      # another object with equal content: not the cached one
    assert core.mask3(m2, 3, torch.float32) is not c
    assert core.mask3(None, 3, torch.float32) is None
    with torch.inference_mode():                                        # inference tensors track no version: never cached
        mi = torch.ones(2, 4, dtype=torch.long)
        d1, d2 = core.mask3(mi, 4, torch.float32), core.mask3(mi, 4, torch.float32)
        assert d1 is not d2 and torch.equal(d1, d2)


def test_param_grads_one_conversion_for_the_block_same_values_as_one_per_row():
    g = torch.randn(5, 64)
    metas = [(torch.bfloat16, (1, 1, 64)), (torch.bfloat16, (64,)), None, (torch.bfloat16, (4, 16)), (torch.bfloat16, (64,))]
    out = fused._param_grads(g, metas)
    assert out[2] is None
    for i, m in enumerate(metas):
        if m is not None:
            assert out[i].dtype == m[0] and tuple(out[i].shape) == tuple(m[1])
            assert torch.equal(out[i].reshape(-1), g[i].to(m[0]))
    # rows of one block share a base (ONE conversion kernel on the GPU) ...
    assert out[0]._base is not None and out[0]._base is out[1]._base
    # ... unless the parameters' dtypes differ: then every row is converted on its own
    mixed = fused._param_grads(g[:2], [(torch.bfloat16, (64,)), (torch.float32, (64,))])
    assert mixed[0].dtype == torch.bfloat16 and mixed[1].dtype == torch.float32
    assert torch.equal(mixed[1], g[1])
