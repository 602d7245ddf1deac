"""CPU oracle for the RWKV-7 (x070) time-mix / channel-mix math AROUND the WKV op.

TEST INFRASTRUCTURE ONLY (same rule as wkv7_oracle.py).  float64 restatement of

* decode step of time-mix  -> /root/reference/model/llm/rwkv_s2s_single_ffn.py:482-506
* decode step of chan-mix  -> /root/reference/model/llm/rwkv_s2s_single_ffn.py:545-549
* training time-mix        -> /root/reference/model/llm/rwkv_s2s_single_ffn.py:158-196
* training channel-mix     -> /root/reference/model/llm/rwkv_s2s_single_ffn.py:223-230
* block                    -> /root/reference/model/llm/rwkv_s2s_single_ffn.py:251-259

Every WKV evaluation goes through ``wkv7_oracle.wkv7_forward`` (the op oracle), so the
golden fixtures made by tests/golden/make_golden.py pin both files at once.
Weight names are the BlinkDL ones the reference uses (fla names map via
/root/reference/utils/convert_rwkv.py:17-41).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .wkv7_oracle import wkv7_forward

F64 = torch.float64


def _d(x):
    return x.detach().to("cpu", F64)


def _wkv_inputs(xr, xw, xk, xv, xa, xg, W, layer_id, v_first, H, N):
    """Shared by the step and sequence forms: projections, LoRAs, kk, k-update, v-residual.
    x* are [..., C]; returns r, w_pre (BlinkDL pre-activation), k, v, kk, a_gate, g, v_first."""
    r = xr @ W["R_"]
    wl = torch.tanh(xw @ W["w1"]) @ W["w2"]
    k = xk @ W["K_"]
    v = xv @ W["V_"]
    a = torch.sigmoid(W["a0"] + (xa @ W["a1"]) @ W["a2"])
    g = torch.sigmoid(xg @ W["g1"]) @ W["g2"]
    kk = F.normalize((k * W["k_k"]).reshape(*k.shape[:-1], H, N), dim=-1, p=2.0).reshape(k.shape)
    k = k * (1 + (a - 1) * W["k_a"])
    if layer_id == 0:
        v_first = v
    else:
        v = v + (v_first - v) * torch.sigmoid(W["v0"] + (xv @ W["v1"]) @ W["v2"])
    # decay = exp(-0.606531*sigmoid(w0+wl))  ==  exp(-exp(w_pre)),  w_pre = -softplus(-(w0+wl)) - 0.5
    w_pre = -F.softplus(-(W["w0"] + wl)) - 0.5
    return r, w_pre, k, v, kk, a, g, v_first


def tmix_seq(layer_id, H, N, x, x_prev, v_first, state, W):
    """T steps of RWKV_x070_TMix_one (:482-506) == RWKV_x070_TMix_seq (:509-540).
    x [T,C]; x_prev [C]; v_first [T,C]; state [H,N,N] value-major.
    Returns out [T,C], x_prev_T, state_T, v_first_out."""
    W = {n: _d(t) for n, t in W.items()}
    x, x_prev, v_first, state = _d(x), _d(x_prev), _d(v_first), _d(state)
    T, C = x.shape
    xx = torch.cat((x_prev[None], x[:-1])) - x                          # :483 / :511
    mix = lambda m: x + xx * W[m]
    r, w_pre, k, v, kk, a, g, v_first = _wkv_inputs(
        mix("x_r"), mix("x_w"), mix("x_k"), mix("x_v"), mix("x_a"), mix("x_g"), W, layer_id, v_first, H, N)
    sh = lambda t: t.reshape(1, T, H, N)
    y, sT = wkv7_forward(sh(w_pre), sh(r), sh(k), sh(v), sh(-kk), sh(kk * a), s0=state[None])   # :497-502
    y = y.reshape(T, C)
    y = F.group_norm(y, num_groups=H, weight=W["ln_w"], bias=W["ln_b"], eps=64e-5)                 # :504
    y = y + ((r * k * W["r_k"]).reshape(T, H, N).sum(-1, keepdim=True) * v.reshape(T, H, N)).reshape(T, C)
    return (y * g) @ W["O_"], x[-1], sT[0], v_first                                               # :506


def cmix_seq(x, x_prev, x_k, K_, V_):
    """T steps of RWKV_x070_CMix_one (:545-549)."""
    x, x_prev, x_k, K_, V_ = map(_d, (x, x_prev, x_k, K_, V_))
    xx = torch.cat((x_prev[None], x[:-1])) - x
    k = torch.relu((x + xx * x_k) @ K_) ** 2
    return k @ V_, x[-1]


def tmix_train(sd, prefix, layer_id, H, N, x, mask, v_first, eps):
    """RWKV_Tmix_x070.forward (:158-196) from a state_dict with BlinkDL names.
    x [B,T,C]; mask [B,T,1]."""
    g_ = lambda n: _d(sd[prefix + n])
    x, mask = _d(x), _d(mask)
    B, T, C = x.shape
    x = x * mask                                                         # :160
    xx = F.pad(x, (0, 0, 1, -1)) - x                                     # :162 ZeroPad2d((0,0,1,-1))
    mix = lambda n: x + xx * g_(n)
    lin = lambda n, t: t @ g_(n + ".weight").T
    xr, xw, xk, xv, xa, xg = (mix(n) for n in ("x_r", "x_w", "x_k", "x_v", "x_a", "x_g"))
    r = lin("receptance", xr)
    w = -F.softplus(-(g_("w0") + torch.tanh(xw @ g_("w1")) @ g_("w2"))) - 0.5   # :172
    k = lin("key", xk)
    v = lin("value", xv)
    r, w, k, v = r * mask, w * mask, k * mask, v * mask                  # :175-178
    if layer_id == 0:
        v_first = v
    else:
        v = v + (_d(v_first) - v) * torch.sigmoid(g_("v0") + (xv @ g_("v1")) @ g_("v2"))
    a = torch.sigmoid(g_("a0") + (xa @ g_("a1")) @ g_("a2"))
    g = torch.sigmoid(xg @ g_("g1")) @ g_("g2")
    kk = F.normalize((k * g_("k_k")).view(B, T, H, N), dim=-1, p=2.0).view(B, T, C) * mask   # :186-188
    k = k * (1 + (a - 1) * g_("k_a"))
    v = v * mask
    sh = lambda t: t.reshape(B, T, H, N)
    y, _ = wkv7_forward(sh(w), sh(r), sh(k), sh(v), sh(-kk), sh(kk * a))                     # :191
    y = y.reshape(B * T, C)
    y = F.group_norm(y, H, g_("ln_x.weight"), g_("ln_x.bias"), eps=eps).view(B, T, C)       # :192
    y = y + ((sh(r) * sh(k) * g_("r_k")).sum(-1, keepdim=True) * sh(v)).view(B, T, C)       # :194
    return lin("output", y * g), v_first


def cmix_train(sd, prefix, x, mask):
    """RWKV_CMix_x070.forward (:223-230)."""
    x = _d(x) * _d(mask)
    xx = F.pad(x, (0, 0, 1, -1)) - x
    k = x + xx * _d(sd[prefix + "x_k"])
    k = torch.relu(k @ _d(sd[prefix + "key.weight"]).T) ** 2
    return k @ _d(sd[prefix + "value.weight"]).T


def block_train(sd, layer_id, H, N, x, mask, v_first, head_size_divisor=8):
    """Block.forward (:251-259)."""
    x = _d(x)
    C = x.shape[-1]
    ln = lambda n, t: F.layer_norm(t, (C,), _d(sd[n + ".weight"]), _d(sd[n + ".bias"]))
    if layer_id == 0:
        x = ln("ln0", x)
    att, v_first = tmix_train(sd, "att.", layer_id, H, N, ln("ln1", x), mask, v_first,
                              eps=1e-5 * head_size_divisor ** 2)
    x = x + att
    x = x + cmix_train(sd, "ffn.", ln("ln2", x), mask)
    return x, v_first
