"""Per-kernel parity: every C-ABI entry point against the oracle's math on seeded inputs.
GPU only (-m gpu).  Float tolerances are written next to each check; integer outputs are exact."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from _util import rel_err  # noqa: E402


def dev():
    return torch.device('cuda:0')


@pytest.fixture(scope='module')
def ops():
    from plankassembly_b200 import ops as _ops
    return _ops


# ------------------------------------------------------------------------------ embeddings
@pytest.mark.parametrize('n_tables,d', [(5, 512), (4, 128), (5, 256)])
def test_embed_input_fwd_bwd(ops, n_tables, d):
    g = torch.Generator().manual_seed(0)
    rows = [514, 300, 4, 3, 2][:n_tables]
    B, S = 3, 77
    ids = [torch.randint(0, r, (B, S), generator=g) for r in rows]
    tabs = [torch.randn(r, d, generator=g, requires_grad=True) for r in rows]
    ref = sum(t[i] for t, i in zip(tabs, ids))
    w = torch.randn(B, S, d, generator=g)
    (ref * w).sum().backward()
    ctabs = [t.detach().to(dev()).requires_grad_(True) for t in tabs]
    out = ops.EmbedInput.apply(n_tables, False, *[i.to(dev()) for i in ids], *ctabs)
    assert rel_err(out.cpu(), ref.detach()) < 1e-6
    (out * w.to(dev())).sum().backward()
    for a, b in zip(ctabs, tabs):
        assert rel_err(a.grad.cpu(), b.grad) < 1e-5


def test_embed_output_fwd_bwd(ops):
    g = torch.Generator().manual_seed(1)
    B, T, d, dof = 5, 64, 128, 6
    value = torch.randint(0, 514, (B, T), generator=g)
    ev = torch.randn(514, d, generator=g, requires_grad=True)
    ec = torch.randn(dof, d, generator=g, requires_grad=True)
    ep = torch.randn(math.ceil(T / dof), d, generator=g, requires_grad=True)
    t = torch.arange(T - 1)
    ref = torch.cat([torch.zeros(B, 1, d), ev[value[:, :-1]] + ec[t % dof][None] + ep[t // dof][None]], 1)
    w = torch.randn(B, T, d, generator=g)
    (ref * w).sum().backward()
    cv, cc, cp = (x.detach().to(dev()).requires_grad_(True) for x in (ev, ec, ep))
    out = ops.EmbedOutput.apply(value.to(dev()), T, dof, False, cv, cc, cp)
    assert torch.equal(out.cpu()[:, 0], torch.zeros(B, d))
    assert rel_err(out.cpu(), ref.detach()) < 1e-6
    (out * w.to(dev())).sum().backward()
    for a, b in ((cv, ev), (cc, ec), (cp, ep)):
        assert rel_err(a.grad.cpu(), b.grad) < 1e-5


# ------------------------------------------------------------------------------ residual + LN
@pytest.mark.parametrize('d,eps,with_a', [(512, 1.0, True), (128, 1.0, True), (256, 1e-5, False), (512, 1e-5, False)])
def test_add_ln_fwd_bwd(ops, d, eps, with_a):
    g = torch.Generator().manual_seed(2)
    rows = 301
    x = torch.randn(rows, d, generator=g, requires_grad=True)
    a = torch.randn(rows, d, generator=g, requires_grad=True) if with_a else None
    gamma = (1 + 0.1 * torch.randn(d, generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(d, generator=g)).requires_grad_(True)
    ref = F.layer_norm(x + a if with_a else x, (d,), gamma, beta, eps)
    w = torch.randn(rows, d, generator=g)
    (ref * w).sum().backward()
    cx, cg, cb = (t.detach().to(dev()).requires_grad_(True) for t in (x, gamma, beta))
    ca = a.detach().to(dev()).requires_grad_(True) if with_a else None
    out = ops.AddLayerNorm.apply(cx, ca, None, cg, cb, eps, 0.0)
    assert rel_err(out.cpu(), ref.detach()) < 2e-6
    (out * w.to(dev())).sum().backward()
    assert rel_err(cx.grad.cpu(), x.grad) < 1e-5
    if with_a:
        assert rel_err(ca.grad.cpu(), a.grad) < 1e-5
    assert rel_err(cg.grad.cpu(), gamma.grad) < 1e-5
    assert rel_err(cb.grad.cpu(), beta.grad) < 1e-5


def test_add_ln_dropout_mask_consistent(ops):
    """Dropout inside the fused LN: keep-rate ~ 1-p, scale 1/(1-p), and bwd regenerates the fwd mask."""
    g = torch.Generator().manual_seed(3)
    rows, d, p = 512, 512, 0.2
    x = torch.zeros(rows, d, device=dev(), requires_grad=True)
    a = torch.ones(rows, d, device=dev(), requires_grad=True)
    gamma = torch.ones(d, device=dev(), requires_grad=True)
    beta = torch.zeros(d, device=dev(), requires_grad=True)
    # x = 0, a = 1: the pre-norm row is dropout(a) in {0, 1/(1-p)}; kept entries lie above the row mean, dropped ones below
    y = ops.AddLayerNorm.apply(x, a, None, gamma, beta, 1.0, p)
    keep = (y > 0)
    assert abs(keep.float().mean().item() - (1 - p)) < 5e-3
    ref = torch.nn.functional.layer_norm(keep.float() / (1 - p), (d,), eps=1.0)
    assert torch.allclose(y, ref, atol=1e-6)
    w = torch.randn(rows, d, generator=g).to(dev())
    (y * w).sum().backward()
    assert torch.equal(a.grad != 0, keep & (x.grad != 0))
    assert torch.allclose(a.grad[keep], x.grad[keep] / (1 - p), rtol=1e-6, atol=0)


def test_relu_dropout(ops):
    g = torch.Generator().manual_seed(4)
    z = torch.randn(64, 1024, generator=g)
    cz = z.to(dev()).requires_grad_(True)
    out = ops.ReluDropout.apply(cz * 1.0, 0.0)
    assert torch.equal(out.cpu(), torch.relu(z))
    out.sum().backward()
    assert torch.equal(cz.grad.cpu(), (z > 0).float())
    p = 0.2
    out = ops.ReluDropout.apply(cz * 1.0, p)
    pos = (z > 0).to(dev())
    keep = out > 0
    assert abs(keep[pos].float().mean().item() - (1 - p)) < 1e-2
    assert torch.allclose(out[keep], cz.detach()[keep] / (1 - p))


# ------------------------------------------------------------------------------ attention
def attn_ref(q, k, v, kpm, causal, H):
    """float64 oracle math (plank_oracle._mha core) on [B,L,d] tensors."""
    B, Lq, d = q.shape
    Lk, dh = k.shape[1], d // H
    qh, kh, vh = (t.view(B, -1, H, dh).transpose(1, 2) for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2) / math.sqrt(dh)
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :], float('-inf'))
    if causal:
        s = s + torch.triu(torch.full((Lq, Lk), float('-inf'), dtype=s.dtype), 1)
    return (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, Lq, d)


SELF_CASES = [  # B, H, dh, L, causal, masked-tail
    (2, 4, 32, 299, False, True), (2, 8, 64, 512, False, True), (3, 4, 32, 64, True, True),
    (2, 8, 64, 256, True, True), (1, 2, 32, 1, True, False), (1, 8, 32, 130, False, False), (2, 8, 64, 70, True, False),
]


@pytest.mark.parametrize('B,H,dh,L,causal,tail', SELF_CASES)
def test_self_attention_fwd_bwd(ops, B, H, dh, L, causal, tail):
    g = torch.Generator().manual_seed(5)
    d = H * dh
    qkv = torch.randn(B, L, 3 * d, generator=g, dtype=torch.float64, requires_grad=True)
    kpm = torch.zeros(B, L, dtype=torch.bool)
    if tail:
        for b in range(B):
            kpm[b, max(1, L - 1 - 17 * (b + 1)):] = True
    q, k, v = qkv.split(d, -1)
    ref = attn_ref(q, k, v, kpm if tail else None, causal, H)
    w = torch.randn(B, L, d, generator=g, dtype=torch.float64)
    (ref * w).sum().backward()
    cq = qkv.detach().float().to(dev()).requires_grad_(True)
    ck = kpm.to(dev()).view(torch.uint8) if tail else None
    out = ops.SelfAttention.apply(cq, None, ck, H, causal, 0.0, 0)
    assert rel_err(out.cpu(), ref.detach()) < 5e-6
    (out * w.float().to(dev())).sum().backward()
    assert rel_err(cq.grad.cpu(), qkv.grad) < 2e-5


@pytest.mark.parametrize('B,H,dh,Lq,Lk', [(2, 4, 32, 64, 299), (2, 8, 64, 256, 512), (1, 8, 64, 3, 1199), (2, 4, 32, 128, 5)])
def test_cross_attention_fwd_bwd(ops, B, H, dh, Lq, Lk):
    g = torch.Generator().manual_seed(6)
    d = H * dh
    q = torch.randn(B, Lq, d, generator=g, dtype=torch.float64, requires_grad=True)
    kv = torch.randn(B, Lk, 2 * d, generator=g, dtype=torch.float64, requires_grad=True)
    kpm = torch.zeros(B, Lk, dtype=torch.bool)
    for b in range(B):
        kpm[b, max(1, Lk - 3 - 40 * b):] = True
    k, v = kv.split(d, -1)
    ref = attn_ref(q, k, v, kpm, False, H)
    w = torch.randn(B, Lq, d, generator=g, dtype=torch.float64)
    (ref * w).sum().backward()
    cq = q.detach().float().to(dev()).requires_grad_(True)
    ckv = kv.detach().float().to(dev()).requires_grad_(True)
    out = ops.CrossAttention.apply(cq, ckv, None, kpm.to(dev()).view(torch.uint8), H, 0.0, 0)
    assert rel_err(out.cpu(), ref.detach()) < 5e-6
    (out * w.float().to(dev())).sum().backward()
    assert rel_err(cq.grad.cpu(), q.grad) < 2e-5
    assert rel_err(ckv.grad.cpu(), kv.grad) < 2e-5


def test_attention_dropout_fwd_bwd_consistent(ops):
    """With V = identity the output IS the dropped probability matrix, which exposes the mask; the
    backward must then equal autograd through softmax * mask / (1-p) with that same mask."""
    g = torch.Generator().manual_seed(7)
    B, H, dh, L, p = 2, 2, 64, 64, 0.2
    d = H * dh
    q = torch.randn(B, L, d, generator=g)
    k = torch.randn(B, L, d, generator=g)
    v = torch.eye(L)[None, :, None, :].expand(B, L, H, dh).reshape(B, L, d).contiguous()
    qkv = torch.cat([q, k, v], -1)
    cq = qkv.to(dev()).requires_grad_(True)
    out = ops.SelfAttention.apply(cq, None, None, H, False, p, 0)
    pd = out.detach().cpu().view(B, L, H, dh).transpose(1, 2)            # [B,H,L,L] dropped probs
    q64 = qkv.double().requires_grad_(True)
    qq, kk, vv = q64.split(d, -1)
    qh, kh, vh = (t.view(B, L, H, dh).transpose(1, 2) for t in (qq, kk, vv))
    prob = torch.softmax(qh @ kh.transpose(-1, -2) / math.sqrt(dh), -1)
    mask = (pd > 0).double()
    assert abs(mask.mean().item() - (1 - p)) < 2e-2
    ref = ((prob * mask / (1 - p)) @ vh).transpose(1, 2).reshape(B, L, d)
    assert rel_err(out.cpu(), ref.detach()) < 5e-6
    w = torch.randn(B, L, d, generator=g)
    (ref * w.double()).sum().backward()
    (out * w.to(dev())).sum().backward()
    assert rel_err(cq.grad.cpu(), q64.grad) < 2e-5


# ------------------------------------------------------------------------------ heads
def dist_train_ref(lv, lp_raw, sw, d, eps=1e-6):
    T = lv.shape[1]
    upper = torch.triu(torch.ones(T, T, dtype=torch.bool))
    lp = (lp_raw / d).masked_fill(upper[None], eps)
    pi = torch.sigmoid(sw)[..., None]
    return torch.cat([torch.log_softmax(lv, -1) + torch.log(torch.clamp(1 - pi, min=eps)),
                      torch.log_softmax(lp, -1) + torch.log(torch.clamp(pi, min=eps))], -1)


@pytest.mark.parametrize('B,T,V', [(4, 64, 514), (2, 256, 514), (3, 7, 514)])
def test_dist_loss_fwd_bwd(ops, B, T, V):
    g = torch.Generator().manual_seed(8)
    d, PAD = 512, 513
    lv = (2 * torch.randn(B, T, V, generator=g, dtype=torch.float64)).requires_grad_(True)
    lp = (300 * torch.randn(B, T, T, generator=g, dtype=torch.float64)).requires_grad_(True)
    sw = torch.randn(B, T, generator=g, dtype=torch.float64).requires_grad_(True)
    label = torch.randint(0, 512, (B, T), generator=g)
    for b in range(B):
        for i in range(1, T):
            if torch.rand(1, generator=g).item() < 0.3:
                label[b, i] = V + int(torch.randint(0, i, (1,), generator=g))
        n = T - 1 - 2 * b
        if n >= 1:
            label[b, n:] = PAD
    dists = dist_train_ref(lv, lp, sw, d)
    valid = label != PAD
    picked = dists.gather(-1, label.clamp(max=V + T - 1)[..., None])[..., 0]
    loss = -(picked * valid).sum() / valid.sum()
    acc = ((dists.argmax(-1) == label) & valid).sum() / valid.sum()
    loss.backward()
    c = [t.detach().float().to(dev()).requires_grad_(True) for t in (lv, lp, sw)]
    closs, cacc, cpred = ops.DistLoss.apply(c[0], c[1], c[2], label.to(dev()), PAD, 1.0 / d)
    assert abs(closs.item() - loss.item()) < 2e-5 * abs(loss.item())
    assert abs(cacc.item() - acc.item()) < 1e-6
    assert (cpred.cpu() == dists.argmax(-1)).float().mean() > 0.999        # fp32-vs-fp64 near-ties only
    full = ops.dist_train_full(c[0].detach(), c[1].detach(), c[2].detach(), 1.0 / d)
    assert rel_err(full.cpu(), dists.detach()) < 1e-5
    (closs * 3.0).backward()
    assert rel_err(c[0].grad.cpu(), 3 * lv.grad) < 2e-5
    assert rel_err(c[1].grad.cpu(), 3 * lp.grad) < 2e-5
    assert rel_err(c[2].grad.cpu(), 3 * sw.grad) < 2e-5


# ------------------------------------------------------------------------------ decode pieces
def test_decode_attn_matches_oracle():
    from plankassembly_b200._lib import call
    g = torch.Generator().manual_seed(9)
    for (B, H, dh, Tmax, t) in [(3, 4, 32, 64, 0), (3, 4, 32, 64, 37), (2, 8, 64, 256, 255)]:
        d = H * dh
        kc = torch.randn(B, Tmax, d, generator=g)
        vc = torch.randn(B, Tmax, d, generator=g)
        qkv = torch.randn(B, 3 * d, generator=g)
        ck, cv, cq = kc.to(dev()), vc.to(dev()), qkv.to(dev())
        o = torch.empty(B, d, device=dev())
        base = cq.data_ptr()
        call('pa_decode_attn', base, 3 * d, base + 4 * d, base + 8 * d, 3 * d, ck.data_ptr(), cv.data_ptr(), Tmax, d, t, t + 1,
             None, None, None, B, H, dh, dh ** -0.5, o.data_ptr(), torch.cuda.current_stream().cuda_stream)
        kc[:, t], vc[:, t] = qkv[:, d:2 * d], qkv[:, 2 * d:]
        assert torch.equal(ck.cpu()[:, t], kc[:, t]) and torch.equal(cv.cpu()[:, t], vc[:, t])
        ref = attn_ref(qkv[:, None, :d].double(), kc[:, :t + 1].double(), vc[:, :t + 1].double(), None, False, H)[:, 0]
        assert rel_err(o.cpu(), ref) < 5e-6
    # cross attention against a packed [B,S,2d] projection with a key-padding mask
    B, H, dh, S = 2, 8, 64, 299
    d = H * dh
    kv = torch.randn(B, S, 2 * d, generator=g)
    q = torch.randn(B, d, generator=g)
    kpm = torch.zeros(B, S, dtype=torch.bool)
    kpm[0, 200:] = True
    kpm[1, 17:] = True
    ckv, cq, ckpm = kv.to(dev()), q.to(dev()), kpm.to(dev()).view(torch.uint8)
    o = torch.empty(B, d, device=dev())
    call('pa_decode_attn', cq.data_ptr(), d, None, None, 0, ckv.data_ptr(), ckv.data_ptr() + 4 * d, S, 2 * d, 0, S,
         None, ckpm.data_ptr(), None, B, H, dh, dh ** -0.5, o.data_ptr(), torch.cuda.current_stream().cuda_stream)
    ref = attn_ref(q[:, None].double(), kv[..., :d].double(), kv[..., d:].double(), kpm, False, H)[:, 0]
    assert rel_err(o.cpu(), ref) < 5e-6
    # with kv_len (1 + last non-PAD key) the PAD tail is not streamed at all: same result, bit for bit
    from plankassembly_b200 import ops as _ops
    kv_len = _ops._kv_len(ckpm)
    assert kv_len.cpu().tolist() == [200, 17]
    o2 = torch.empty(B, d, device=dev())
    call('pa_decode_attn', cq.data_ptr(), d, None, None, 0, ckv.data_ptr(), ckv.data_ptr() + 4 * d, S, 2 * d, 0, S,
         None, ckpm.data_ptr(), kv_len.data_ptr(), B, H, dh, dh ** -0.5, o2.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert torch.equal(o2, o)


def test_add_ln_folded_linear_bias(ops):
    """y = LN(x + a + bias): the bias of the producing linear folded into the LN kernels; its gradient
    (column sums of da) must come out of the backward kernel."""
    g = torch.Generator().manual_seed(31)
    rows, d = 515, 512
    x = torch.randn(rows, d, generator=g, requires_grad=True)
    a = torch.randn(rows, d, generator=g, requires_grad=True)
    bias = torch.randn(d, generator=g, requires_grad=True)
    gamma = (1 + 0.1 * torch.randn(d, generator=g)).requires_grad_(True)
    beta = (0.1 * torch.randn(d, generator=g)).requires_grad_(True)
    ref = F.layer_norm(x + a + bias, (d,), gamma, beta, 1.0)
    w = torch.randn(rows, d, generator=g)
    (ref * w).sum().backward()
    c = [t.detach().to(dev()).requires_grad_(True) for t in (x, a, bias, gamma, beta)]
    out = ops.AddLayerNorm.apply(c[0], c[1], c[2], c[3], c[4], 1.0, 0.0)
    assert rel_err(out.cpu(), ref.detach()) < 2e-6
    (out * w.to(dev())).sum().backward()
    for ours, theirs in zip(c, (x, a, bias, gamma, beta)):
        assert rel_err(ours.grad.cpu(), theirs.grad) < 1e-5


def test_attention_bwd_bias_gradient(ops):
    """The tensor-core backward kernels also emit the in-projection bias gradient (column sums of dqkv)."""
    g = torch.Generator().manual_seed(32)
    B, H, dh, L = 2, 8, 64, 200
    d = H * dh
    qkv = torch.randn(B, L, 3 * d, generator=g).to(dev()).requires_grad_(True)
    bias = torch.zeros(3 * d, device=dev(), requires_grad=True)
    w = torch.randn(B, L, d, generator=g).to(dev())
    out = ops.SelfAttention.apply(qkv, bias, None, H, True, 0.0, 1)
    (out * w).sum().backward()
    assert rel_err(bias.grad.cpu(), qkv.grad.sum((0, 1)).cpu()) < 1e-5
    q = torch.randn(B, 70, d, generator=g).to(dev()).requires_grad_(True)
    kv = torch.randn(B, L, 2 * d, generator=g).to(dev()).requires_grad_(True)
    bias2 = torch.zeros(3 * d, device=dev(), requires_grad=True)
    out = ops.CrossAttention.apply(q, kv, bias2, None, H, 0.0, 1)
    (out * torch.randn(B, 70, d, generator=g).to(dev())).sum().backward()
    ref = torch.cat([q.grad.sum((0, 1)), kv.grad.sum((0, 1))])
    assert rel_err(bias2.grad.cpu(), ref.cpu()) < 1e-5


def test_gemm_skinny_f32_exact():
    """Decode projections: fp32 FMA GEMM against fp64; must be at fp32 rounding level (not TF32)."""
    from plankassembly_b200._lib import call
    g = torch.Generator().manual_seed(33)
    for M, N, K, relu in [(64, 1536, 512, False), (64, 512, 1024, False), (64, 1024, 512, True), (4, 514, 128, False), (7, 1, 256, False), (130, 384, 128, False)]:
        x, w, b = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
        cx, cw, cb = x.to(dev()), w.to(dev()), b.to(dev())
        out = torch.full((M, N), float('nan'), device=dev())
        call('pa_gemm_skinny_f32', cx.data_ptr(), K, cw.data_ptr(), K, cb.data_ptr(), out.data_ptr(), N, M, N, K, int(relu),
             torch.cuda.current_stream().cuda_stream)
        ref = x.double() @ w.double().T + b.double()
        if relu:
            ref = torch.relu(ref)
        assert rel_err(out.cpu(), ref) < 2e-6, (M, N, K)
