"""GPU parity tests of the fused time-mix kernels (rwkvtts_b200/fused.py -> csrc/tmix_fused.cu) against a plain
PyTorch fp32 restatement of the same reference lines (rwkv_s2s_single_ffn.py:160-195, :226), forward and backward.

Tolerance: outputs / activation gradients are bf16 (rounding alone ~1.6e-3 relative-L2) -> 5e-3; per-channel
parameter gradients are fp32 sums of products of bf16 tensors -> 1e-2.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
ACT_TOL, PAR_TOL = 5e-3, 1e-2


def rel(a, b):
    a, b = a.float(), b.float()
    return float(((a - b).norm() / b.norm().clamp_min(1e-20)).detach())


@pytest.fixture(scope="module")
def fused():
    import rwkvtts_b200  # noqa: F401
    from rwkvtts_b200 import fused
    assert torch.cuda.is_available()
    return fused


def _mask(B, T, dev):
    m = torch.ones(B, T, 1, device=dev, dtype=torch.bfloat16)
    for b in range(B):
        m[b, T - 1 - (b * 3) % max(T // 2, 1):] = 0
    return m


def ref_shift_mix(x, mixes, mask, prev):
    x = x.float()
    if mask is not None:
        x = x * mask.float()
    first = torch.zeros_like(x[:, :1]) if prev is None else prev.float().unsqueeze(1)
    xx = torch.cat((first, x[:, :-1]), dim=1) - x
    return [x + xx * m.float() for m in mixes]


@pytest.mark.parametrize("B,T,C,n,use_mask,use_prev", [(2, 37, 192, 6, False, False), (2, 64, 1024, 6, True, False),
                                                       (1, 5, 2048, 1, False, True), (3, 16, 768, 1, True, True),
                                                       (1, 1, 64, 6, False, True),
                                                       # the four-channel adjoint (C >= 512, >= 64 rows): masks, carried
                                                       # shift state, every width class, uneven rows per CTA
                                                       (3, 333, 1024, 6, True, True), (2, 70, 768, 6, False, True),
                                                       (5, 41, 2048, 6, True, False), (4, 128, 512, 1, True, True),
                                                       (2, 100, 1024, 1, False, False)])
def test_shift_mix(fused, B, T, C, n, use_mask, use_prev):
    g = torch.Generator(device="cuda").manual_seed(B * 100 + T)
    x = torch.randn(B, T, C, device="cuda", generator=g).bfloat16().requires_grad_(True)
    mixes = [torch.rand(1, 1, C, device="cuda", generator=g).bfloat16().requires_grad_(True) for _ in range(n)]
    mask = _mask(B, T, "cuda") if use_mask else None
    prev = torch.randn(B, C, device="cuda", generator=g).bfloat16() if use_prev else None
    outs = fused.shift_mix(x, mixes, mask, prev)
    douts = [torch.randn(B, T, C, device="cuda", generator=g).bfloat16() for _ in range(n)]
    torch.autograd.backward(outs, douts)
    xr = x.detach().clone().requires_grad_(True)
    mr = [m.detach().float().requires_grad_(True) for m in mixes]
    refs = ref_shift_mix(xr, mr, mask, prev)
    torch.autograd.backward(refs, [d.float() for d in douts])
    for o, r in zip(outs, refs):
        assert rel(o, r) < ACT_TOL
    assert rel(x.grad, xr.grad) < ACT_TOL
    for m, r in zip(mixes, mr):
        assert rel(m.grad, r.grad) < PAR_TOL


def ref_prep(k, v, w_lo, a_lo, v_lo, v_first, w0, a0, v0, k_k, k_a, mask, H):
    f = lambda t: None if t is None else t.float()
    k, v, w_lo, a_lo, v_lo, v_first = map(f, (k, v, w_lo, a_lo, v_lo, v_first))
    B, T, C = k.shape
    w = -F.softplus(-(w0.float() + w_lo)) - 0.5
    if mask is not None:
        m = mask.float()
        w, k, v = w * m, k * m, v * m
    if v_lo is not None:
        v = v + (v_first - v) * torch.sigmoid(v0.float() + v_lo)
    a = torch.sigmoid(a0.float() + a_lo)
    kk = F.normalize((k * k_k.float()).view(B, T, H, 64), dim=-1, p=2.0).view(B, T, C)
    if mask is not None:
        kk, v = kk * m, v * m
    k2 = k * (1 + (a - 1) * k_a.float())
    return w, k2, v, -kk, kk * a


@pytest.mark.parametrize("B,T,C,has_v,use_mask", [(2, 33, 192, True, False), (2, 64, 1024, True, True),
                                                  (1, 16, 768, False, False), (2, 9, 128, False, True)])
def test_prep(fused, B, T, C, has_v, use_mask):
    H = C // 64
    g = torch.Generator(device="cuda").manual_seed(C + T)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()
    acts = [rn(B, T, C) for _ in range(6)]                       # k, v, w_lo, a_lo, v_lo, v_first
    if not has_v:
        acts[4] = acts[5] = None
    pars = [(0.5 * rn(1, 1, C)) for _ in range(5)]               # w0, a0, v0, k_k, k_a
    pars[3] = pars[3] + 1
    if not has_v:
        pars[2] = None
    mask = _mask(B, T, "cuda") if use_mask else None
    leaf = lambda t: None if t is None else t.detach().clone().requires_grad_(True)
    a1, p1 = [leaf(t) for t in acts], [leaf(t) for t in pars]
    outs = fused.prep(*a1, p1[0], p1[1], p1[2], p1[3], p1[4], mask)
    douts = [rn(B, T, C) for _ in range(5)]
    torch.autograd.backward(outs, douts)
    a2, p2 = [leaf(t) for t in acts], [None if t is None else t.detach().float().requires_grad_(True) for t in pars]
    refs = ref_prep(*a2, p2[0], p2[1], p2[2], p2[3], p2[4], mask, H)
    torch.autograd.backward(refs, [d.float() for d in douts])
    for name, o, r in zip(("w", "k2", "v2", "a_op", "b_op"), outs, refs):
        assert rel(o, r) < ACT_TOL, name
    for name, x1, x2 in zip(("k", "v", "w_lo", "a_lo", "v_lo", "v_first"), a1, a2):
        if x1 is not None:
            assert rel(x1.grad, x2.grad) < 2 * ACT_TOL, name
    for name, x1, x2 in zip(("w0", "a0", "v0", "k_k", "k_a"), p1, p2):
        if x1 is not None:
            assert rel(x1.grad, x2.grad) < PAR_TOL, name


def ref_out(y, r, k2, v2, g, r_k, ln_w, ln_b, eps, H):
    y, r, k2, v2, g = (t.float() for t in (y, r, k2, v2, g))
    B, T, C = y.shape
    z = F.group_norm(y.reshape(B * T, C), H, ln_w.float(), ln_b.float(), eps).view(B, T, C)
    z = z + ((r.view(B, T, H, 64) * k2.view(B, T, H, 64) * r_k.float().view(H, 64)).sum(-1, keepdim=True)
             * v2.view(B, T, H, 64)).view(B, T, C)
    return z * g


@pytest.mark.parametrize("B,T,C", [(2, 33, 192), (2, 64, 1024), (1, 7, 2048)])
def test_out(fused, B, T, C):
    H = C // 64
    g_ = torch.Generator(device="cuda").manual_seed(C)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g_).bfloat16()
    acts = [rn(B, T, C) for _ in range(5)]
    pars = [0.3 * rn(H, 64), 1 + 0.3 * rn(C), 0.3 * rn(C)]
    leaf = lambda t: t.detach().clone().requires_grad_(True)
    a1, p1 = [leaf(t) for t in acts], [leaf(t) for t in pars]
    o = fused.out(*a1, *p1, 64e-5)
    d_o = rn(B, T, C)
    o.backward(d_o)
    a2, p2 = [leaf(t) for t in acts], [t.detach().float().requires_grad_(True) for t in pars]
    ref = ref_out(*a2, *p2, 64e-5, H)
    ref.backward(d_o.float())
    assert rel(o, ref) < ACT_TOL
    for name, x1, x2 in zip(("y", "r", "k2", "v2", "g"), a1, a2):
        assert rel(x1.grad, x2.grad) < 2 * ACT_TOL, name
    for name, x1, x2 in zip(("r_k", "ln_w", "ln_b"), p1, p2):
        assert rel(x1.grad, x2.grad) < PAR_TOL, name


@pytest.mark.parametrize("B,T,C,use_res,use_bias", [(2, 33, 256, True, True), (8, 64, 1024, True, True), (1, 7, 2048, False, True),
                                                    (3, 5, 512, True, False),
                                                    # warp-per-row adjoint: every C / 256 class, odd row counts
                                                    (3, 333, 1024, True, True), (2, 77, 768, True, False), (1, 9, 256, False, True),
                                                    (5, 41, 512, False, False),
                                                    # ring adjoint at its widest row (3 CTAs per SM by shared memory)
                                                    (2, 96, 2048, True, True)])
def test_add_layernorm(fused, B, T, C, use_res, use_bias):
    g = torch.Generator(device="cuda").manual_seed(C + T)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g).bfloat16()
    x, res = rn(B, T, C), (rn(B, T, C) if use_res else None)
    w, b = 1 + 0.3 * rn(C), (0.3 * rn(C) if use_bias else None)
    dy, ds = rn(B, T, C), rn(B, T, C)
    leaf = lambda t: None if t is None else t.detach().clone().requires_grad_(True)
    x1, r1, w1, b1 = leaf(x), leaf(res), leaf(w), leaf(b)
    y, tot = fused.add_layernorm(x1, r1, w1, b1, 1e-5)
    torch.autograd.backward([y, tot], [dy, ds])
    f = lambda t: None if t is None else t.detach().float().requires_grad_(True)
    x2, r2, w2, b2 = f(x), f(res), f(w), f(b)
    tot2 = x2 if r2 is None else (x2 + r2).bfloat16().float() + 0 * (x2 + r2)      # the sum is a bf16 tensor in the model
    tot2 = x2 if r2 is None else x2 + r2
    y2 = F.layer_norm(tot2, (C,), w2, b2, 1e-5)
    torch.autograd.backward([y2, tot2], [dy.float(), ds.float()])
    assert rel(y, y2) < ACT_TOL and rel(tot, tot2) < ACT_TOL
    assert rel(x1.grad, x2.grad) < 2 * ACT_TOL
    if use_res:
        assert rel(r1.grad, r2.grad) < 2 * ACT_TOL
    assert rel(w1.grad, w2.grad) < PAR_TOL
    if use_bias:
        assert rel(b1.grad, b2.grad) < PAR_TOL
    with torch.no_grad():                          # inference path (no stats)
        y3, _ = fused.add_layernorm(x, res, w, b, 1e-5)
    assert torch.equal(y3, y)


def test_sqrelu(fused):
    x = torch.randn(3, 17, 512, device="cuda").bfloat16().requires_grad_(True)
    dy = torch.randn(3, 17, 512, device="cuda").bfloat16()
    y = fused.sqrelu(x)
    y.backward(dy)
    xr = x.detach().clone().requires_grad_(True)
    yr = torch.relu(xr) ** 2
    yr.backward(dy)
    assert torch.equal(y, yr)                      # same roundings as ATen: relu is exact, one rounding in the square
    assert rel(x.grad, xr.grad) < ACT_TOL


@pytest.mark.parametrize("layer_id,use_mask,mask_rwk", [(0, False, True), (1, False, True), (1, True, True),
                                                        (1, True, False), (0, True, False)])
def test_tmix_fused_matches_aten_path(fused, layer_id, use_mask, mask_rwk):
    """core.tmix with the fused kernels against the same function with the ATen chain (same weights, same WKV op)."""
    from rwkvtts_b200 import core
    from rwkvtts_b200.x070 import RWKV_Tmix_x070
    import types
    args = types.SimpleNamespace(n_embd=256, dim_att=256, n_layer=4, head_size_a=64, head_size_divisor=8)
    torch.manual_seed(3)
    mod = RWKV_Tmix_x070(args, layer_id).cuda().bfloat16()
    with torch.no_grad():
        for p in mod.parameters():                 # the reference zero-inits several tensors: make every path matter
            if float(p.abs().sum()) == 0:
                p.copy_(0.1 * torch.randn_like(p))
    B, T = 2, 48
    x = torch.randn(B, T, 256, device="cuda").bfloat16()
    vf = torch.randn(B, T, 256, device="cuda").bfloat16() if layer_id else None
    mask = _mask(B, T, "cuda") if use_mask else None
    dout = torch.randn(B, T, 256, device="cuda").bfloat16()
    res = {}
    for flag in (True, False):
        core.FUSED = flag
        mod.zero_grad(set_to_none=True)
        xi = x.clone().requires_grad_(True)
        vfi = None if vf is None else vf.clone().requires_grad_(True)
        out, v_first, _, _ = core.tmix(mod.params(), layer_id, xi, vfi, mask, mask_rwk=mask_rwk)
        out.backward(dout)
        if layer_id == 0:
            res.setdefault("vf", []).append(v_first.detach())
        res[flag] = (out.detach(), xi.grad, None if vfi is None else vfi.grad,
                     {n: p.grad.clone() for n, p in mod.named_parameters() if p.grad is not None})
    core.FUSED = True
    if layer_id == 0:
        assert rel(res["vf"][0], res["vf"][1]) < ACT_TOL
    assert rel(res[True][0], res[False][0]) < 2e-2
    assert rel(res[True][1], res[False][1]) < 5e-2      # both sides are bf16 chains; the ATen one rounds after every op
    if vf is not None:
        assert rel(res[True][2], res[False][2]) < 5e-2
    for n, gr in res[False][3].items():
        assert n in res[True][3], n
        # wiring check: the per-kernel tests above pin the math against fp32; the ATen chain rounds every intermediate
        # and every partial parameter gradient to bf16, so the two sides differ by a few per cent on small gradients
        assert rel(res[True][3][n], gr) < 0.15, (n, rel(res[True][3][n], gr))


def test_wide_rows_take_the_aten_chain_under_autograd_and_the_abi_refuses_them(fused):
    """the adjoint kernels cannot launch rows wider than 2048 channels (C/8 threads per row in 256-thread
    CTAs).  C = 2560: usable() is false under autograd (the ATen chain trains), true without it; the *_backward entry points
    answer RWKVTTS_ERR_SHAPE instead of failing in the launch."""
    from rwkvtts_b200 import _lib
    C = 2560
    x = torch.randn(2, 8, C, device="cuda").bfloat16()
    assert not fused.usable(x) and not fused.ln_usable(x)
    with torch.no_grad():
        assert fused.usable(x) and fused.ln_usable(x)
        mixes = [torch.rand(C, device="cuda").bfloat16() for _ in range(6)]
        outs = fused.shift_mix(x, mixes)                               # forward kernels do cover it
        want = ref_shift_mix(x, mixes, None, None)
        for o, w in zip(outs, want):
            assert rel(o, w) < ACT_TOL
    L = _lib.lib()
    f32 = torch.zeros(6 * C, device="cuda")
    import ctypes
    dptr = (ctypes.c_void_p * 6)(*[x.data_ptr()] * 6)
    rc = L.rwkvtts_tmix_shift_mix_backward(2, 8, C, 6, x.data_ptr(), None, None, f32.data_ptr(), dptr, x.data_ptr(),
                                           f32.data_ptr(), f32.data_ptr(), None)
    assert rc == -1                                                    # RWKVTTS_ERR_SHAPE
    # a whole time-mix layer at that width trains (ATen chain around the WKV kernels)
    from rwkvfla.layers.rwkv7 import RWKV7Attention
    att = RWKV7Attention(hidden_size=C, head_dim=64, layer_idx=1, num_hidden_layers=2).cuda().to(torch.bfloat16)
    h = torch.randn(1, 32, C, device="cuda").bfloat16().requires_grad_(True)
    out, _, _, _ = att(h, v_first=torch.randn(1, 32, C, device="cuda").bfloat16())
    out.float().square().mean().backward()
    assert torch.isfinite(h.grad.float()).all()
