"""`rwkvfla.modules.l2warp.l2_warp` (train_rwkv7_asr_jsonl.py:27): identity on the loss whose backward
adds a pull of the largest logit towards 0 (same as L2Wrap, rwkv_s2s_single_ffn.py:261-274)."""
import torch


class _L2Wrap(torch.autograd.Function):
    @staticmethod
    def forward(ctx, loss, logits, l2_penalty_factor):
        ctx.save_for_backward(logits)
        ctx.factor = l2_penalty_factor
        return loss

    @staticmethod
    def backward(ctx, grad_output):
        (logits,) = ctx.saved_tensors
        factor = ctx.factor / (logits.shape[0] * logits.shape[1])
        maxx, ids = torch.max(logits, -1, keepdim=True)
        glogits = torch.zeros_like(logits)
        glogits.scatter_(-1, ids, maxx * factor)
        return grad_output, glogits, None


def l2_warp(loss, logits, l2_penalty_factor: float = 1e-4):
    return _L2Wrap.apply(loss, logits, l2_penalty_factor)
