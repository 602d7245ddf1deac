"""CPU oracle for the WKV-7 state recurrence (forward, backward, stateful step).

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``rwkvtts_b200``) may import
this module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker.

It restates, in float64 torch on the CPU, the algorithm of the reference's native ops:

* forward recurrence      -> /root/reference/model/llm/cuda/wkv7_cuda.cu:17-42
* backward (adjoint)      -> /root/reference/model/llm/cuda/wkv7_cuda.cu:62-129
* stateful forward / step -> /root/reference/model/llm/cuda/rwkv7_state_fwd_fp16.cu:18-56
                             (same step, state loaded from / written back to ``state``)
* the same step in PyTorch-> /root/reference/model/llm/rwkv_s2s_single_ffn.py:497-502

Parity pin: the reference holds no golden vectors for this path, so
the oracle is pinned against outputs of the reference's own Python run in the build
container (``tests/golden/make_golden.py`` imports ``rwkv_s2s_single_ffn.py`` with the
CUDA loader stubbed and records ``RWKV_x070_TMix_one`` outputs) and cross-checked against
the pure-torch ``dplr_recurrence`` of the installed flash-linear-attention 0.5.1.

Conventions (reference op boundary, wkv7_op.cpp:7): argument order (w, q, k, v, a, b)
where the reference names the last two ``z`` and ``a``; tensors are [B, T, H, C];
state is value-major ``S[b, h, i=value, j=key]``; ``w`` is the BlinkDL pre-activation,
decay = exp(-exp(w)).
"""
from __future__ import annotations

import torch

F64 = torch.float64


def _f64(*xs):
    return [x.detach().to("cpu", F64) for x in xs]


def wkv7_forward(w, q, k, v, a, b, s0=None, return_states=False):
    """y_t = S_t q_t with S_t = S_{t-1} diag(d_t) + (S_{t-1} a_t) b_t^T + v_t k_t^T.

    Follows wkv7_cuda.cu:17-42 (decay :21, sa :27-31, update :39, output :40).
    Returns (y [B,T,H,C] f64, S_T [B,H,C,C] f64[, states [B,T+1,H,C,C]]).
    """
    w, q, k, v, a, b = _f64(w, q, k, v, a, b)
    B, T, H, C = w.shape
    S = torch.zeros(B, H, C, C, dtype=F64) if s0 is None else s0.detach().to("cpu", F64).clone()
    d = torch.exp(-torch.exp(w))
    y = torch.empty(B, T, H, C, dtype=F64)
    states = [S.clone()] if return_states else None
    for t in range(T):
        sa = torch.einsum("bhij,bhj->bhi", S, a[:, t])
        S = S * d[:, t, :, None, :] + sa[..., None] * b[:, t, :, None, :] \
            + v[:, t, :, :, None] * k[:, t, :, None, :]
        y[:, t] = torch.einsum("bhij,bhj->bhi", S, q[:, t])
        if return_states:
            states.append(S.clone())
    if return_states:
        return y, S, torch.stack(states, dim=1)
    return y, S


def wkv7_backward(w, q, k, v, a, b, dy, s0=None, dsT=None):
    """Explicit adjoint of :func:`wkv7_forward`, restating wkv7_cuda.cu:62-129 in f64.

    The kernel un-steps the state from 16-step snapshots (:76-94); here every S_{t-1} is
    kept from a forward pass instead, which is the same quantity.  Gradient names follow
    the forward arguments.  ``dw`` carries the chain through d = exp(-exp(w)) (:108).
    Returns (dw, dq, dk, dv, da, db, ds0).
    """
    w, q, k, v, a, b, dy = _f64(w, q, k, v, a, b, dy)
    B, T, H, C = w.shape
    _, _, states = wkv7_forward(w, q, k, v, a, b, s0=s0, return_states=True)
    d = torch.exp(-torch.exp(w))
    dS = torch.zeros(B, H, C, C, dtype=F64) if dsT is None else dsT.detach().to("cpu", F64).clone()
    dw, dq, dk, dv, da, db = [torch.empty(B, T, H, C, dtype=F64) for _ in range(6)]
    for t in range(T - 1, -1, -1):
        Sp, St = states[:, t], states[:, t + 1]          # S_{t-1}, S_t
        # dq_i-key = sum_value S_t[value][key] dy[value]              (:84-89)
        dq[:, t] = torch.einsum("bhij,bhi->bhj", St, dy[:, t])
        # dS_t += dy_t q_t^T                                            (:95-96)
        dS = dS + dy[:, t, :, :, None] * q[:, t, :, None, :]
        sa = torch.einsum("bhij,bhj->bhi", Sp, a[:, t])
        dd = torch.einsum("bhij,bhij->bhj", dS, Sp)                  # (:103)
        dw[:, t] = dd * d[:, t] * (-torch.exp(w[:, t]))              # (:108)
        dk[:, t] = torch.einsum("bhij,bhi->bhj", dS, v[:, t])        # (:104,:109)
        dv[:, t] = torch.einsum("bhij,bhj->bhi", dS, k[:, t])        # (:105,:110)
        dSb = torch.einsum("bhij,bhj->bhi", dS, b[:, t])             # (:106)
        db[:, t] = torch.einsum("bhij,bhi->bhj", dS, sa)             # (:107,:111)
        da[:, t] = torch.einsum("bhij,bhi->bhj", Sp, dSb)            # (:113-122)
        # dS_{t-1} = dS_t diag(d_t) + dSb a_t^T                        (:125-128)
        dS = dS * d[:, t, :, None, :] + dSb[..., None] * a[:, t, :, None, :]
    return dw, dq, dk, dv, da, db, dS


def wkv7_state_forward(state, r, w, k, v, a, b):
    """Stateful forward (rwkv7_state_fwd_fp16.cu:9-57 / wkv7s.cu:9-57).

    ``state`` [B,H,C,C] value-major; r,w,k,v,a,b [B,T,H*C].  Returns (y [B,T,H*C] f64,
    new_state f64).  With T == 1 this is the per-token decode op
    (rwkv_asr_cuda_whisper.py:702-703).
    """
    B, T, HC = r.shape
    H = state.shape[1]
    C = HC // H
    rs = [x.reshape(B, T, H, C) for x in (w, r, k, v, a, b)]
    y, sT = wkv7_forward(*rs, s0=state)
    return y.reshape(B, T, HC), sT


from rwkvtts_b200.synth import make_inputs  # noqa: E402,F401  (shared input recipe)


def rel_l2(x, ref):
    x = x.detach().to("cpu", F64)
    ref = ref.detach().to("cpu", F64)
    return float((x - ref).norm() / ref.norm().clamp_min(1e-300))


def excess_rel_l2(x_bf16, ref_f64):
    """Compute error of a bf16 result beyond the unavoidable output quantisation.

    rounding y to bf16 alone costs ~1.6e-3 relative-L2, above the 1e-3 bar the north star
    states, so the bar is applied to the error in quadrature beyond that floor:
    sqrt(max(0, err^2 - floor^2)) where floor = rel_l2(bf16(ref), ref).
    Returns (excess, err, floor).
    """
    err = rel_l2(x_bf16.float(), ref_f64)
    floor = rel_l2(ref_f64.to(torch.bfloat16).to(F64), ref_f64)
    exc = max(0.0, err * err - floor * floor) ** 0.5
    return exc, err, floor
