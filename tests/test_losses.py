"""rwkvtts_b200.losses.LabelSmoothingLoss against the reference's class (third_party/cosyvoice/transformer/
label_smoothing_loss.py:20-96) in fp32: value and gradient, with / without smoothing, both normalisations, ignored
positions.  Build container only."""
import importlib.util
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/third_party/cosyvoice/transformer/label_smoothing_loss.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted")
sys.path.insert(0, ROOT)


def test_label_smoothing_loss_matches_reference_value_and_gradient():
    from rwkvtts_b200.losses import LabelSmoothingLoss
    spec = importlib.util.spec_from_file_location("_ref_lsl", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(0)
    for smoothing in (0.0, 0.1, 0.35):
        for norm in (True, False):
            B, T, V = 3, 17, 53
            x = (torch.randn(B, T, V, generator=g) * 2).requires_grad_(True)
            y = torch.randint(0, V, (B, T), generator=g)
            y[0, :5] = -1
            y[2, 9:] = -1
            a = ref.LabelSmoothingLoss(size=V, padding_idx=-1, smoothing=smoothing, normalize_length=norm)(x, y)
            (ga,) = torch.autograd.grad(a, x)
            b = LabelSmoothingLoss(size=V, padding_idx=-1, smoothing=smoothing, normalize_length=norm)(x, y)
            (gb,) = torch.autograd.grad(b, x)
            assert abs(float(a.detach()) - float(b.detach())) < 2e-6 * max(1.0, abs(float(a.detach()))), (smoothing, norm)
            assert float((ga - gb).abs().max()) < 1e-6 and float(gb[0, :5].abs().sum()) == 0.0
    # smoothing 0 is the conventional cross entropy over the valid tokens
    x = torch.randn(2, 5, 11, generator=g)
    y = torch.randint(0, 11, (2, 5), generator=g); y[1, 3:] = -1
    ce = torch.nn.functional.cross_entropy(x.view(-1, 11), y.view(-1), ignore_index=-1)
    assert abs(float(LabelSmoothingLoss(11, -1, 0.0, True)(x, y)) - float(ce)) < 1e-6


def test_xy_channel_losses_equal_the_reference_loss_loop():
    """xy_llm.py:233-240: eight heads, eight nn.CrossEntropyLoss, summed; value and gradients (hidden states, head weights
    and biases) from the logit-free path."""
    from rwkvtts_b200.losses import xy_channel_losses
    g = torch.Generator().manual_seed(1)
    B, T, D = 2, 13, 16
    sizes = [97] + [23] * 7
    heads = torch.nn.ModuleList([torch.nn.Linear(D, v) for v in sizes])
    for lsm in (0.0, 0.1):
        h = torch.randn(B, T, D, generator=g, requires_grad=True)
        labels = torch.stack([torch.randint(0, v, (B, T), generator=g) for v in sizes], dim=-1)
        labels[0, :4, :] = -100
        labels[1, 7:, 3] = -100
        crits = [torch.nn.CrossEntropyLoss(label_smoothing=lsm) for _ in sizes]
        want = 0
        for i in range(8):                                            # the reference's loop
            logits = heads[i](h)
            want = want + crits[i](logits.view(-1, logits.shape[-1]), labels[:, :, i].reshape(-1))
        params = [h] + [p for hd in heads for p in hd.parameters()]
        gw = torch.autograd.grad(want, params)
        got = xy_channel_losses(h, heads, labels, label_smoothing=lsm, num_chunks=3)
        gg = torch.autograd.grad(got, params)
        assert abs(float(want.detach()) - float(got.detach())) < 1e-5
        assert max(float((a - b).abs().max()) for a, b in zip(gw, gg)) < 1e-5
