"""Cross-GPU sequence split of the WKV-7 recurrence for contexts that do not fit one GPU (SURVEY.md section 8 row f4,
section 5 "long context"): every rank holds a contiguous slice of T of the SAME sequences, and the only thing that
crosses NVLink is the recurrent state -- [B, H, 64, 64] fp32, 16 KB per (batch, head) -- once per direction:

    forward :  rank r receives S from rank r-1 (zeros on rank 0), runs its slice from that state with the chunked
               kernels (rwkvtts_wkv7_forward_ex: initial / final state), sends its final state to rank r+1;
    backward:  rank r receives dS from rank r+1 (zeros on the last rank), runs the adjoint of its slice
               (rwkvtts_wkv7_backward_ex: dsT in, ds0 out), sends ds0 to rank r-1.

The recurrence is sequential in T, so the ranks form a pipeline; `head_groups` > 1 cuts the heads into independent
groups that travel through it back to back (rank r works on group g while rank r+1 works on group g-1), which keeps
(G) / (G + N - 1) of the machine busy instead of 1 / N.  The token shift of the time-mix needs one more row per rank:
`shift_boundary()` hands the last token of a slice to the next rank.  The reference has no such path (its contexts fit
one GPU: BASELINE configs stop at T = 8192); the state hand-off is what rwkvfla's chunk_rwkv7(initial_state,
output_final_state) offers and what SURVEY section 5 sketches.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist

from . import ops


def _neighbours(group):
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    g = lambda r: dist.get_global_rank(group, r) if group is not None else r
    return rank, world, (g(rank - 1) if rank > 0 else None), (g(rank + 1) if rank + 1 < world else None)


class _SeqParWkv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, w, q, k, v, a, b, group):
        rank, world, prev, nxt = _neighbours(group)
        B, T, H, C = w.shape
        s0 = None
        if prev is not None:
            s0 = torch.empty(B, H, C, C, dtype=torch.float32, device=w.device)
            dist.recv(s0, src=prev, group=group)
        y = torch.empty_like(v)
        sT = torch.empty(B, H, C, C, dtype=torch.float32, device=w.device)
        train = any(ctx.needs_input_grad[:6])
        if train:
            s = torch.empty(B, H, T // ops.CHUNK_LEN, C, C, dtype=torch.float32, device=w.device)
            sa = torch.empty(B, T, H, C, dtype=torch.float32, device=w.device)
            ops.wkv7_forward_(w, q, k, v, a, b, y, s, sa, s0=s0, sT=sT)
            ctx.save_for_backward(w, q, k, v, a, b, s, sa, s0, sT)
        else:
            ops.wkv7_forward_infer_(w, q, k, v, a, b, y, s0=s0, sT=sT)
        if nxt is not None:
            dist.send(sT, dst=nxt, group=group)
        ctx.group = group
        return y

    @staticmethod
    def backward(ctx, dy):
        w, q, k, v, a, b, s, sa, s0, sT = ctx.saved_tensors
        rank, world, prev, nxt = _neighbours(ctx.group)
        dsT = None
        if nxt is not None:
            dsT = torch.empty_like(sT)
            dist.recv(dsT, src=nxt, group=ctx.group)
        grads = [torch.empty_like(x) for x in (w, q, k, v, a, b)]
        ds0 = torch.empty_like(sT) if prev is not None else None
        ops.wkv7_backward_(w, q, k, v, a, b, dy.contiguous(), s, sa, *grads, s0=s0, dsT=dsT, ds0=ds0, sT=sT)
        if prev is not None:
            dist.send(ds0, dst=prev, group=ctx.group)
        return (*grads, None)


def wkv7_sequence_parallel(w, q, k, v, a, b, group=None, head_groups: int = 1):
    """y of this rank's slice.  w,q,k,v,a,b: bf16 [B, T_local, H, 64] contiguous (op order), T_local % 16 == 0, the
    slices of the ranks of `group` in rank order form the full sequences.  Differentiable."""
    H = w.shape[2]
    if head_groups <= 1 or H % head_groups != 0:
        return _SeqParWkv.apply(w, q, k, v, a, b, group)
    hs = H // head_groups
    outs = []
    for g in range(head_groups):
        sl = [t[:, :, g * hs:(g + 1) * hs].contiguous() for t in (w, q, k, v, a, b)]
        outs.append(_SeqParWkv.apply(*sl, group))
    return torch.cat(outs, dim=2)


class _ShiftBoundary(torch.autograd.Function):
    @staticmethod
    def forward(ctx, last_row, group):
        rank, world, prev, nxt = _neighbours(group)
        got = torch.zeros_like(last_row)
        # even ranks send first, odd ranks receive first: no cycle in the blocking pairs
        ops_ = []
        if nxt is not None:
            ops_.append(dist.P2POp(dist.isend, last_row.contiguous(), nxt, group))
        if prev is not None:
            ops_.append(dist.P2POp(dist.irecv, got, prev, group))
        for r in (dist.batch_isend_irecv(ops_) if ops_ else []):
            r.wait()
        ctx.group = group
        return got

    @staticmethod
    def backward(ctx, d_got):
        rank, world, prev, nxt = _neighbours(ctx.group)
        d_last = torch.zeros_like(d_got)
        ops_ = []
        if prev is not None:
            ops_.append(dist.P2POp(dist.isend, d_got.contiguous(), prev, ctx.group))
        if nxt is not None:
            ops_.append(dist.P2POp(dist.irecv, d_last, nxt, ctx.group))
        for r in (dist.batch_isend_irecv(ops_) if ops_ else []):
            r.wait()
        return d_last, None


def shift_boundary(last_row: torch.Tensor, group=None) -> torch.Tensor:
    """Token-shift state across the split: gives every rank the last row [B, C] of the previous rank's slice (zeros on
    rank 0) -- the `prev` argument of fused.shift_mix / core.tmix(shift_state=) -- and routes its gradient back."""
    return _ShiftBoundary.apply(last_row, group)
