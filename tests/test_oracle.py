"""Pins the oracle (oracle/*.py, oracle/wkv7_oracle.c) -- CPU only.

Pins, in order of strength:
 1. golden fixtures produced by the reference's own Python (tests/golden/make_golden.py);
 2. autograd through the forward loop (gradient oracle, SURVEY.md section 8 "Operator definition");
 3. the pure-torch dplr_recurrence of the installed flash-linear-attention (third-party stand-in
    for rwkvfla; state is key-major there, so it is compared after a transpose);
 4. the C port against the f64 oracle.
"""
import os

import pytest
import torch

from oracle import rwkv7_model_oracle as MO
from oracle import wkv7_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
ORDER = "wqkvab"


@pytest.mark.parametrize("layer_id", [0, 1])
def test_golden_tmix_one(layer_id):
    g = torch.load(f"{GOLD}/tmix_one_L{layer_id}.pt")
    out, x_prev, state, vf = MO.tmix_seq(layer_id, g["H"], g["N"], g["x"], g["x_prev0"], g["v_first_in"],
                                         g["state0"], g["weights"])
    # reference ran in fp32: agreement to fp32 round-off over 24 chained steps
    assert O.rel_l2(out, g["out"]) < 2e-6
    assert O.rel_l2(state, g["state_T"]) < 2e-6          # value-major layout, same orientation
    assert O.rel_l2(vf, g["v_first_out"]) < 2e-6
    assert torch.equal(x_prev.float(), g["x_prev_T"])


def test_golden_cmix_one():
    g = torch.load(f"{GOLD}/cmix_one.pt")
    out, x_prev = MO.cmix_seq(g["x"], g["x_prev0"], g["x_k"], g["K_"], g["V_"])
    assert O.rel_l2(out, g["out"]) < 2e-6
    assert torch.equal(x_prev.float(), g["x_prev_T"])


@pytest.mark.parametrize("layer_id", [0, 1])
def test_golden_block(layer_id):
    g = torch.load(f"{GOLD}/block_L{layer_id}.pt")
    a = g["args"]
    H = a["n_embd"] // a["head_size_a"]
    y, vf = MO.block_train(g["state_dict"], layer_id, H, a["head_size_a"], g["x"], g["mask"], g["v_first_in"],
                           head_size_divisor=a["head_size_divisor"])
    assert O.rel_l2(y, g["y"]) < 5e-6
    assert O.rel_l2(vf, g["v_first_out"]) < 5e-6


def _autograd_forward(w, q, k, v, a, b, s0):
    B, T, H, C = w.shape
    S = s0
    d = torch.exp(-torch.exp(w))
    ys = []
    for t in range(T):
        sa = torch.einsum("bhij,bhj->bhi", S, a[:, t])
        S = S * d[:, t, :, None, :] + sa[..., None] * b[:, t, :, None, :] + v[:, t, :, :, None] * k[:, t, :, None, :]
        ys.append(torch.einsum("bhij,bhj->bhi", S, q[:, t]))
    return torch.stack(ys, 1), S


def test_backward_is_adjoint_of_forward():
    x = O.make_inputs(2, 40, 3, seed=1, dtype=torch.float64)
    s0 = torch.randn(2, 3, 64, 64, dtype=torch.float64) * 0.1
    dsT = torch.randn(2, 3, 64, 64, dtype=torch.float64) * 0.1
    leaves = [x[n].clone().requires_grad_(True) for n in ORDER]
    s0l = s0.clone().requires_grad_(True)
    y, sT = _autograd_forward(*leaves, s0l)
    (y * x["dy"]).sum().add((sT * dsT).sum()).backward()
    grads = O.wkv7_backward(*[x[n] for n in ORDER], x["dy"], s0=s0, dsT=dsT)
    for n, leaf, g in zip(ORDER, leaves, grads[:6]):
        assert (leaf.grad - g).abs().max() < 1e-11, n
    assert (s0l.grad - grads[6]).abs().max() < 1e-11


def test_finite_difference_dw():
    x = O.make_inputs(1, 8, 1, seed=3, dtype=torch.float64)
    args = [x[n] for n in ORDER]
    dw = O.wkv7_backward(*args, x["dy"])[0]
    eps = 1e-6
    for idx in [(0, 0, 0, 5), (0, 3, 0, 17), (0, 7, 0, 63)]:
        wp, wm = x["w"].clone(), x["w"].clone()
        wp[idx] += eps
        wm[idx] -= eps
        fp = (O.wkv7_forward(wp, *args[1:])[0] * x["dy"]).sum()
        fm = (O.wkv7_forward(wm, *args[1:])[0] * x["dy"]).sum()
        assert abs((fp - fm) / (2 * eps) - dw[idx]) < 1e-6 * max(1.0, abs(float(dw[idx])))


def test_state_forward_chains():
    """Running T steps at once == running them in two stateful calls (a4/a5 semantics)."""
    x = O.make_inputs(2, 24, 2, seed=5, dtype=torch.float64)
    flat = {n: t.reshape(2, 24, 128) for n, t in x.items()}
    s0 = torch.randn(2, 2, 64, 64, dtype=torch.float64) * 0.1
    y, sT = O.wkv7_state_forward(s0, flat["q"], flat["w"], flat["k"], flat["v"], flat["a"], flat["b"])
    cut = 9
    h = lambda n, sl: flat[n][:, sl]
    y1, s1 = O.wkv7_state_forward(s0, *[h(n, slice(0, cut)) for n in "qwkvab"])
    y2, s2 = O.wkv7_state_forward(s1, *[h(n, slice(cut, None)) for n in "qwkvab"])
    assert torch.allclose(torch.cat([y1, y2], 1), y, atol=1e-12)
    assert torch.allclose(s2, sT, atol=1e-12)


def test_against_fla_dplr_recurrence():
    naive = pytest.importorskip("fla.ops.generalized_delta_rule.dplr.naive")
    x = O.make_inputs(1, 48, 2, seed=7, dtype=torch.float32)
    y, sT = O.wkv7_forward(*[x[n] for n in ORDER])
    tr = lambda t: t.transpose(1, 2).contiguous()          # fla naive wants [B,H,T,D]
    gk = -torch.exp(x["w"].float())                         # log-decay convention of rwkvfla
    o, fs = naive.dplr_recurrence(tr(x["q"]) * 8.0, tr(x["k"]), tr(x["v"]), tr(x["a"]), tr(x["b"]), tr(gk),
                                  None, True)               # q*sqrt(64) cancels its d_k^-0.5 (naive.py:23)
    assert O.rel_l2(tr(o), y) < 1e-5
    assert O.rel_l2(fs.transpose(-1, -2), sT) < 1e-5        # fla state is [K,V]; ours value-major [V,K]


def test_c_port_matches_f64(c_oracle):
    x = O.make_inputs(2, 64, 2, seed=11)
    args = [x[n] for n in ORDER]
    y64, _ = O.wkv7_forward(*args)
    y, s, sa = c_oracle.c_forward(*args)
    exc, err, floor = O.excess_rel_l2(y, y64)
    assert exc < 1e-5, (exc, err, floor)
    g64 = O.wkv7_backward(*args, x["dy"])
    g = c_oracle.c_backward(*args, x["dy"], s, sa)
    for n, a, b in zip(ORDER, g, g64):
        assert O.excess_rel_l2(a, b)[0] < 1e-4, n
    # snapshot layout: s[b,h,c,j,i] = S[value i][key j] at the end of chunk c (wkv7_cuda.cu:44-50)
    _, _, states = O.wkv7_forward(*args, return_states=True)
    assert O.rel_l2(s[:, :, 1].transpose(-1, -2), states[:, 32]) < 1e-5


def test_c_state_forward(c_oracle):
    x = O.make_inputs(3, 5, 2, seed=13)
    flat = {n: t.reshape(3, 5, 128).contiguous() for n, t in x.items()}
    s0 = (torch.randn(3, 2, 64, 64) * 0.1).contiguous()
    y64, s64 = O.wkv7_state_forward(s0, *[flat[n] for n in "qwkvab"])
    st = s0.clone()
    y = c_oracle.c_state_forward(st, *[flat[n] for n in "qwkvab"])
    assert O.excess_rel_l2(y, y64)[0] < 1e-5
    assert O.rel_l2(st, s64) < 1e-5
