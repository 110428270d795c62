"""Backward-pass kernels (training step, BASELINE config 4) on a B200: each C-ABI entry point against torch autograd of a plain
fp32 statement of the op it differentiates (oracle functions where they exist)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import grounding as og  # noqa: E402


@pytest.fixture(scope="module")
def ops():
    from grove_b200 import ops as _ops
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return _ops


def _rand(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(dtype).cuda()


def _relerr(a, b):
    return float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-12))


# ------------------------------------------------------------------ GEMM extensions
def test_gemm_dact_and_pre_activation(ops):
    M, N, K = 1536, 768, 256
    a, w = _rand((M, K), 1, dtype=torch.bfloat16), _rand((N, K), 2, 1 / math.sqrt(K), dtype=torch.bfloat16)
    bias = _rand((N,), 3)
    ref = a.float() @ w.float().t() + bias
    # forward in training mode: out = GELU(pre), out2 = pre
    o = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    pre = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, o, bias=bias, act="gelu", out2=pre, out2_pre_act=1)
    assert _relerr(pre.float(), ref) < 5e-3 and _relerr(o.float(), F.gelu(ref)) < 6e-3
    # backward: (a @ w^T) * gelu'(pre) and relu mask
    z = _rand((M, N), 4, dtype=torch.bfloat16)
    zz = z.float().requires_grad_(True)
    for kind, fn in (("gelu", F.gelu), ("relu", F.relu)):
        g = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ops.gemm(a, w, g, dact_pre=z, dact=kind)
        (dz,) = torch.autograd.grad(fn(zz).sum(), zz)
        assert _relerr(g.float(), (a.float() @ w.float().t()) * dz) < 6e-3
    # ragged M (rows past the end are clamped for the pre-activation fetch and never stored) with a bias, on the CTA-pair fast epilogue
    Mr = 1400
    g = torch.full((Mr + 8, N), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a[:Mr], w, g[:Mr], bias=bias, dact_pre=z[:Mr].contiguous(), dact="gelu")
    (dz,) = torch.autograd.grad(F.gelu(zz).sum(), zz)
    assert _relerr(g[:Mr].float(), ref[:Mr] * dz[:Mr]) < 6e-3
    assert bool((g[Mr:] == 7.0).all())
    # fp32 output: out2 = value after the activation, before gate and residual
    alpha = torch.tensor([0.4], device="cuda")
    x = _rand((M, N), 5)
    xs = x.clone()
    r = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm(a, w, xs, bias=bias, act="relu", gate_alpha=alpha, resid=xs, out2=r, out2_pre_act=2)
    assert _relerr(r.float(), F.relu(ref)) < 5e-3
    assert _relerr(xs, x + math.tanh(0.4) * F.relu(ref)) < 2e-5


@pytest.mark.parametrize("M,N,K", [(4096, 128, 256), (32768, 128, 256), (8192, 256, 128), (1024, 4096, 4096), (16384, 768, 768)])
def test_wgrad_splitk(ops, M, N, K):
    dy, x = _rand((M, N), 6, dtype=torch.bfloat16), _rand((M, K), 7, dtype=torch.bfloat16)
    out = torch.ones(N, K, device="cuda")
    ops.wgrad(dy, x, out, accumulate=True)
    ref = dy.float().t() @ x.float() + 1.0
    assert _relerr(out, ref) < 3e-5, _relerr(out, ref)


def test_transpose(ops):
    for dt in (torch.float32, torch.bfloat16):
        x = _rand((1000, 384), 8, dtype=dt)
        assert torch.equal(ops.transpose_to_bf16(x), x.to(torch.bfloat16).t().contiguous())


@pytest.mark.parametrize("V,T,G,C,N,kt", [(1, 8, 16, 256, 256, 3), (2, 8, 32, 256, 256, 3), (1, 8, 64, 256, 512, 3), (3, 1, 32, 256, 256, 1),
                                          (1, 1, 64, 512, 256, 1)])
def test_conv_wgrad(ops, V, T, G, C, N, kt):
    tokens = V * T * G * G
    x, dy = _rand((tokens, C), 9, dtype=torch.bfloat16), _rand((tokens, N), 10, dtype=torch.bfloat16)
    taps = 9 * kt
    out = torch.zeros(N, taps * C, device="cuda")
    ops.conv_wgrad(dy, x, out, V=V, T=T, G=G, kt=kt, accumulate=False)
    if kt == 3:
        w = torch.zeros(N, C, 3, 3, 3, device="cuda", requires_grad=True)
        xi = x.float().view(V, T, G, G, C).permute(0, 4, 1, 2, 3)
        y = F.conv3d(xi, w, padding=1)
        (gw,) = torch.autograd.grad(y, w, dy.float().view(V, T, G, G, N).permute(0, 4, 1, 2, 3))
        ref = gw.permute(0, 2, 3, 4, 1).reshape(N, -1)
    else:
        w = torch.zeros(N, C, 3, 3, device="cuda", requires_grad=True)
        xi = x.float().view(V * T, G, G, C).permute(0, 3, 1, 2)
        y = F.conv2d(xi, w, padding=1)
        (gw,) = torch.autograd.grad(y, w, dy.float().view(V * T, G, G, N).permute(0, 3, 1, 2))
        ref = gw.permute(0, 2, 3, 1).reshape(N, -1)
    assert _relerr(out, ref) < 1e-4, _relerr(out, ref)


# ------------------------------------------------------------------ element-wise / reductions
@pytest.mark.parametrize("D", [256, 768, 1280])
def test_layernorm_bwd(ops, D):
    rows = 1000
    x, r, dy = _rand((rows, D), 11), _rand((rows, D), 12), _rand((rows, D), 13)
    gamma, beta = _rand((D,), 14) * 0.3 + 1, _rand((D,), 15)
    dx_in = _rand((rows, D), 16)
    xr = x.clone().requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    br = beta.clone().requires_grad_(True)
    y = og.layer_norm(xr + r, gr, br, 1e-6)
    gx, gg, gb = torch.autograd.grad(y, (xr, gr, br), dy)
    dx = torch.empty_like(x)
    dxb = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    ops.layernorm_bwd(x, gamma, dy, eps=1e-6, r=r, dx_in=dx_in, dx_out=dx, dx_bf16=dxb, dgamma=dg, dbeta=db)
    assert _relerr(dx, gx + dx_in) < 1e-5 and _relerr(dg, gg) < 1e-4 and _relerr(db, gb) < 1e-4
    assert torch.equal(dxb, dx.to(torch.bfloat16))
    # bf16 upstream gradient, frozen affine
    dyb = dy.to(torch.bfloat16)
    ops.layernorm_bwd(x, gamma, dyb, eps=1e-6, r=r, dx_out=dx)
    (gx2,) = torch.autograd.grad(og.layer_norm(xr + r, gamma, beta, 1e-6), xr, dyb.float())
    assert _relerr(dx, gx2) < 1e-5


def test_layernorm_bwd_keys(ops):
    Fr, N, B, C = 3, 256, 5, 256
    keys = _rand((Fr * N, C), 17, dtype=torch.bfloat16)
    src_of = torch.tensor([0, 0, 1, 2, 2], dtype=torch.int32, device="cuda")
    delta, dy = _rand((B * N, C), 18), _rand((B * N, C), 19)
    gamma, beta = _rand((C,), 20) * 0.3 + 1, _rand((C,), 21)
    u = (keys.float().view(Fr, N, C)[src_of.long()].reshape(B * N, C) + delta).requires_grad_(True)
    gr = gamma.clone().requires_grad_(True)
    gx, gg = torch.autograd.grad(og.layer_norm(u, gr, beta, 1e-5), (u, gr), dy)
    dx = torch.empty(B * N, C, device="cuda")
    dg, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    ops.layernorm_bwd(keys, gamma, dy, eps=1e-5, r=delta, dx_out=dx, dgamma=dg, dbeta=db, keys_src_of=src_of, keys_N=N)
    assert _relerr(dx, gx) < 1e-5 and _relerr(dg, gg) < 1e-4 and _relerr(db, dy.sum(0)) < 1e-4


def test_adapter_gate_bwd(ops):
    rows, D = 2048, 768
    conv = _rand((rows, D), 22).requires_grad_(True)
    alpha = torch.tensor([0.37], device="cuda", requires_grad=True)
    dy = _rand((rows, D), 23)
    r = F.relu(conv).to(torch.bfloat16)
    y = torch.tanh(alpha) * F.relu(conv.to(torch.bfloat16).float())
    # reference on the bf16-rounded relu output (what the forward saved)
    rr = r.float().requires_grad_(True)
    y = torch.tanh(alpha) * rr
    g_r, g_a = torch.autograd.grad(y, (rr, alpha), dy)
    ref_dyc = g_r * (r.float() > 0)
    dyc = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    dbias, dalpha = torch.zeros(D, device="cuda"), torch.zeros(1, device="cuda")
    ops.adapter_gate_bwd(dy, r, alpha.detach(), dyc, dbias, dalpha)
    assert _relerr(dyc.float(), ref_dyc) < 5e-3
    assert _relerr(dbias, ref_dyc.sum(0)) < 1e-4 and _relerr(dalpha, g_a) < 1e-4


def test_small_reductions(ops):
    x = _rand((3000, 128), 24, dtype=torch.bfloat16)
    out = torch.ones(128, device="cuda")
    assert _relerr(ops.colsum(x, out), x.float().sum(0) + 1) < 1e-5
    xf = _rand((777, 4), 25)
    assert _relerr(ops.colsum(xf, torch.zeros(4, device="cuda")), xf.sum(0)) < 1e-5
    xs = _rand((7, 1024), 26)
    off = torch.tensor([0, 2, 2, 7], dtype=torch.int32, device="cuda")
    o = ops.segment_sum(xs, off, torch.empty(3, 1024, device="cuda"))
    assert _relerr(o, torch.stack([xs[0:2].sum(0), xs[2:2].sum(0), xs[2:7].sum(0)])) < 1e-6
    dy, xx = _rand((190, 50), 27), _rand((190, 300), 28)
    dw = torch.zeros(50, 300, device="cuda")
    assert _relerr(ops.small_wgrad(dy, xx, dw), dy.t() @ xx) < 1e-5
    y = torch.sigmoid(_rand((100, 4), 29))
    assert _relerr(ops.act_bwd(dy[:100, :4].contiguous(), y, "sigmoid"), dy[:100, :4] * y * (1 - y)) < 1e-6
    x16 = _rand((6, 4096), 30, dtype=torch.bfloat16)
    assert _relerr(ops.batch_sum_bf16(x16, 6), x16.float().sum(0)) < 1e-6


def test_token_self_attention_bwd(ops):
    B, T, H, dh = 9, 6, 8, 32
    q, k, v, do = (_rand((B, T, H * dh), s) for s in (31, 32, 33, 34))
    qq, kk, vv = (t.clone().requires_grad_(True) for t in (q, k, v))
    sep = lambda t: t.view(B, T, H, dh).transpose(1, 2)
    a = torch.softmax(sep(qq) @ sep(kk).transpose(-1, -2) / math.sqrt(dh), -1)
    o = (a @ sep(vv)).transpose(1, 2).reshape(B, T, H * dh)
    rq, rk, rv = torch.autograd.grad(o, (qq, kk, vv), do)
    dq, dk, dv = ops.token_self_attention_bwd(q, k, v, do, B, T, H, dh)
    assert _relerr(dq, rq) < 1e-5 and _relerr(dk, rk) < 1e-5 and _relerr(dv, rv) < 1e-5


def test_decoder_cross_attention_bwd(ops):
    Fr, B, T, N, H, dh = 2, 3, 6, 1024, 8, 16
    src_of = torch.tensor([0, 1, 1], dtype=torch.int32, device="cuda")
    q = _rand((B, T, H * dh), 35)
    k, v = _rand((Fr * N, H * dh), 36, dtype=torch.bfloat16), _rand((Fr * N, H * dh), 37, dtype=torch.bfloat16)
    datt = _rand((B, T, H * dh), 38)
    lse = torch.empty(B, T, H, device="cuda")
    att = ops.t2i_attention(q, k, v, src_of, B, T, N, H, dh, lse=lse)
    qq = q.clone().requires_grad_(True)
    kk = k.float().view(Fr, N, H * dh)[src_of.long()].clone().requires_grad_(True)     # per instance
    vv = v.float().view(Fr, N, H * dh)[src_of.long()].clone().requires_grad_(True)
    sep = lambda t: t.view(B, -1, H, dh).transpose(1, 2)
    a = torch.softmax(sep(qq) @ sep(kk).transpose(-1, -2) / math.sqrt(dh), -1)
    o = (a @ sep(vv)).transpose(1, 2).reshape(B, T, H * dh)
    assert _relerr(att, o) < 1e-4
    rq, rk, rv = torch.autograd.grad(o, (qq, kk, vv), datt)
    dq, dk, dv = ops.t2i_attention_bwd(q, k, v, src_of, att, datt, lse, B, T, N, H, dh)
    assert _relerr(dq, rq) < 2e-4
    assert _relerr(dk.float().view(B, N, -1), rk) < 6e-3 and _relerr(dv.float().view(B, N, -1), rv) < 6e-3
    # image -> token
    qi = _rand((Fr * N, H * dh), 39, dtype=torch.bfloat16)
    kt, vt = _rand((B, T, H * dh), 40), _rand((B, T, H * dh), 41)
    dout = _rand((B * N, H * dh), 42, dtype=torch.bfloat16)
    qq = qi.float().view(Fr, N, H * dh)[src_of.long()].clone().requires_grad_(True)
    kk, vv = kt.clone().requires_grad_(True), vt.clone().requires_grad_(True)
    a = torch.softmax(sep(qq) @ sep(kk).transpose(-1, -2) / math.sqrt(dh), -1)
    o = (a @ sep(vv)).transpose(1, 2).reshape(B, N, H * dh)
    rq, rk, rv = torch.autograd.grad(o, (qq, kk, vv), dout.float().view(B, N, -1))
    dqi, dkt, dvt = ops.i2t_attention_bwd(qi, kt, vt, src_of, dout, B, T, N, H, dh)
    assert _relerr(dqi.float().view(B, N, -1), rq) < 6e-3
    assert _relerr(dkt, rk) < 2e-4 and _relerr(dvt, rv) < 2e-4


# ------------------------------------------------------------------ encoder attention backward
def _ref_attention(qkv, bias, Rh, Rw, heads, hd, G, ws):
    """Attention.forward core (image_encoder.py:304-323) + window_partition/unpartition (:329-384) on a given qkv tensor
    [F,G,G,3,heads,hd]; padded window positions hold the qkv bias (LayerNorm output is zero-padded before the qkv Linear)."""
    Fr = qkv.shape[0]
    if ws:
        Gp = G + (ws - G % ws) % ws
        x = bias.view(1, 1, 1, 3, heads, hd).expand(Fr, Gp, Gp, 3, heads, hd).clone()
        x[:, :G, :G] = qkv
        nw = Gp // ws
        x = x.view(Fr, nw, ws, nw, ws, 3, heads, hd).permute(0, 1, 3, 2, 4, 5, 6, 7).reshape(Fr * nw * nw, ws * ws, 3, heads, hd)
        S = ws
    else:
        x = qkv.reshape(Fr, G * G, 3, heads, hd)
        S = G
    q, k, v = (x[:, :, i].transpose(1, 2) for i in range(3))          # [Bw, heads, S*S, hd]
    attn = (q * hd ** -0.5) @ k.transpose(-2, -1)
    Rh_t, Rw_t = og.rel_pos_table(S, Rh), og.rel_pos_table(S, Rw)
    rq = q.reshape(-1, S, S, hd)
    rel_h = torch.einsum("bhwc,hkc->bhwk", rq, Rh_t)
    rel_w = torch.einsum("bhwc,wkc->bhwk", rq, Rw_t)
    attn = (attn.view(-1, S, S, S, S) + rel_h[..., :, None] + rel_w[..., None, :]).view(-1, heads, S * S, S * S)
    o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(-1, S * S, heads * hd)
    if ws:
        o = o.view(Fr, nw, nw, ws, ws, heads * hd).permute(0, 1, 3, 2, 4, 5).reshape(Fr, Gp, Gp, heads * hd)[:, :G, :G]
    return o.reshape(Fr, G, G, heads * hd)


@pytest.mark.parametrize("G,ws,heads,hd,Fr", [(16, 0, 2, 64, 2), (32, 0, 2, 80, 1), (32, 0, 3, 64, 2), (64, 0, 1, 64, 1), (64, 0, 2, 80, 1), (32, 14, 2, 64, 2),
                                              (64, 14, 2, 80, 1), (16, 14, 3, 64, 1), (64, 14, 3, 64, 2), (32, 14, 2, 80, 2)])
def test_attention_relpos_bwd(ops, G, ws, heads, hd, Fr):
    S = ws if ws else G
    qkv = _rand((Fr, G, G, 3, heads, hd), 43, 0.8, dtype=torch.bfloat16)
    bias = _rand((3 * heads * hd,), 44, 0.3, dtype=torch.bfloat16)
    Rh, Rw = _rand((2 * S - 1, hd), 45, 0.15, dtype=torch.bfloat16), _rand((2 * S - 1, hd), 46, 0.15, dtype=torch.bfloat16)
    datt = _rand((Fr, G, G, heads * hd), 47, dtype=torch.bfloat16)
    x = qkv.float().requires_grad_(True)
    o = _ref_attention(x, bias.float(), Rh.float(), Rw.float(), heads, hd, G, ws)
    (ref,) = torch.autograd.grad(o, x, datt.float())
    att = o.detach().to(torch.bfloat16).contiguous()
    dqkv = torch.full_like(qkv, float("nan"))
    ops.attn_relpos_bwd(qkv, bias if ws else None, Rh, Rw, att, datt, dqkv, F=Fr, G=G, heads=heads, hd=hd, ws=ws)
    torch.cuda.synchronize()
    assert torch.isfinite(dqkv.float()).all()
    for i, name in enumerate("qkv"):
        e = _relerr(dqkv[..., i, :, :].float(), ref[..., i, :, :])
        assert e < 2.5e-2, (name, e)
    # mean error is far below the max-norm bound (bf16 rounding of P and dS)
    assert float((dqkv.float() - ref).abs().mean() / ref.abs().mean()) < 8e-3
    if ws == 0 and G in (32, 64):
        # training path: attention output and row log-sum-exp from the tcgen05 forward kernel -> single-sweep backward kernels
        lse = torch.empty(Fr * G * G, heads, device="cuda")
        att2 = torch.empty_like(att)
        ops.attn_global(qkv, Rh, Rw, att2, F=Fr, G=G, heads=heads, hd=hd, lse=lse)
        dq2 = torch.full_like(qkv, float("nan"))
        ops.attn_relpos_bwd(qkv, None, Rh, Rw, att2, datt, dq2, F=Fr, G=G, heads=heads, hd=hd, ws=0, lse=lse)
        torch.cuda.synchronize()
        assert torch.isfinite(dq2.float()).all()
        for i, name in enumerate("qkv"):
            e = _relerr(dq2[..., i, :, :].float(), ref[..., i, :, :])
            assert e < 2.5e-2, ("fwd-lse path", name, e)
        assert float((dq2.float() - ref).abs().mean() / ref.abs().mean()) < 8e-3
    if ws == 14:
        # training path of the windowed blocks: the tcgen05 window forward saves the row log-sum-exp -> the query-side backward skips its
        # own log-sum-exp sweep
        lse = torch.full((Fr * G * G, heads), float("nan"), device="cuda")
        att2 = torch.empty_like(att)
        tab = ops.window_rel_table(Rh, Rw)
        ops.attn_window_tc(qkv, bias, tab, att2, F=Fr, G=G, heads=heads, hd=hd, ws=14, lse=lse)
        torch.cuda.synchronize()
        assert torch.isfinite(lse).all()
        dq2 = torch.full_like(qkv, float("nan"))
        ops.attn_relpos_bwd(qkv, bias, Rh, Rw, att2, datt, dq2, F=Fr, G=G, heads=heads, hd=hd, ws=14, lse=lse)
        torch.cuda.synchronize()
        assert torch.isfinite(dq2.float()).all()
        for i, name in enumerate("qkv"):
            e = _relerr(dq2[..., i, :, :].float(), ref[..., i, :, :])
            assert e < 2.5e-2, ("window fwd-lse path", name, e)
        assert float((dq2.float() - ref).abs().mean() / ref.abs().mean()) < 8e-3


# ------------------------------------------------------------------ loss derivative
def test_box_losses_bwd(ops):
    B = 64
    g = torch.Generator().manual_seed(48)
    boxes = torch.cat([torch.rand(B, 2, generator=g) * 0.6 + 0.2, torch.rand(B, 2, generator=g) * 0.4 + 0.05], 1).cuda()
    gt = torch.cat([torch.rand(B, 2, generator=g) * 0.6 + 0.2, torch.rand(B, 2, generator=g) * 0.4 + 0.05], 1).cuda()
    gt[:8, :2] = boxes[:8, :2] + 0.45      # some disjoint pairs
    logits = torch.randn(B, generator=g).cuda()
    sel = (torch.rand(B, generator=g) > 0.4).cuda()
    labels = sel.float()
    bb, ll = boxes.clone().requires_grad_(True), logits.clone().requires_grad_(True)
    wg, wo = 2.0, 2.0
    n_gt, n_pred = int(sel.sum()), B
    giou = og.giou_loss_sum(og.box_cxcywh_to_xyxy(bb[sel]), og.box_cxcywh_to_xyxy(gt[sel]))
    l1 = (bb[sel] - gt[sel]).abs().sum()
    bce = F.binary_cross_entropy_with_logits(ll, labels, reduction="sum")
    loss = wg * giou / (n_gt + 1e-8) + wg * l1 / (n_gt + 1e-8) + wo * bce / (n_pred + 1e-8)
    rb, rl = torch.autograd.grad(loss, (bb, ll))
    db, dl = ops.box_losses_bwd(boxes, logits, gt.contiguous(), sel.to(torch.uint8), labels, wg / (n_gt + 1e-8), wo / (n_pred + 1e-8))
    assert _relerr(db, rb) < 1e-4 and _relerr(dl, rl) < 1e-5
