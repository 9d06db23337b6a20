"""tcgen05 TF32 GEMM (pa_gemm_tf32) against fp64 math.  Two references: exact fp64 (tolerance =
TF32 operand rounding, 2^-11 relative per operand) and fp64 on TF32-truncated operands (tight)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from _util import rel_err  # noqa: E402


@pytest.fixture(autouse=True, params=['2', '1', '0'], ids=['umma2sm', 'multicast', 'single'])
def pair_mode(request, monkeypatch):
    """Every case runs in the three CTA organisations of gemm_tc.cu: 2-SM UMMA pairs (cta_group::2, the default), pairs that
    share B through TMA multicast (cta_group::1), single CTAs."""
    monkeypatch.setenv('PLANK_B200_GEMM_PAIR', request.param)
    return request.param


def tf32_trunc(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def check(out, a64, b64, a_t, b_t, tag, alpha=1.0):
    ref = alpha * (a64 @ b64.T)
    ref_t = alpha * (a_t.double() @ b_t.double().T)
    e, et = rel_err(out.cpu(), ref), rel_err(out.cpu(), ref_t)
    print(f'{tag}: rel err vs fp64 {e:.2e}, vs tf32-truncated operands {et:.2e}')
    assert e < 2e-3, tag
    return e, et


@pytest.mark.parametrize('M,N,K', [(256, 128, 128), (1196, 1536, 512), (4096, 514, 512), (300, 1024, 256), (128, 512, 1024), (64, 1536, 512)])
def test_gemm_kmajor_bias(M, N, K):
    from plankassembly_b200 import ops
    g = torch.Generator().manual_seed(0)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g)
    c = torch.full((M, N), float('nan'), device='cuda')
    ops.gemm_tf32(a.cuda(), w.cuda(), c, M, N, K, lda=K, ldb=K, ldc=N, bias=bias.cuda())
    out = c.cpu() - bias
    check(out, a.double(), w.double(), tf32_trunc(a), tf32_trunc(w), f'TN {M}x{N}x{K}')


def test_gemm_relu_dropout_epilogue():
    from plankassembly_b200 import ops
    g = torch.Generator().manual_seed(1)
    M, N, K, p = 512, 1024, 512, 0.2
    a, w, bias = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / K ** 0.5, torch.randn(N, generator=g)
    c = torch.empty(M, N, device='cuda')
    ops.gemm_tf32(a.cuda(), w.cuda(), c, M, N, K, lda=K, ldb=K, ldc=N, bias=bias.cuda(), relu=True)
    ref = torch.relu(a.double() @ w.double().T + bias)
    assert rel_err(c.cpu(), ref) < 2e-3
    c2 = torch.empty(M, N, device='cuda')
    ops.gemm_tf32(a.cuda(), w.cuda(), c2, M, N, K, lda=K, ldb=K, ldc=N, bias=bias.cuda(), relu=True, p_drop=p, seed=7, off=3)
    keep = c2 > 0
    pos = c > 0
    assert abs(keep[pos].float().mean().item() - (1 - p)) < 1e-2
    assert torch.allclose(c2[keep], c[keep] / (1 - p), rtol=1e-6)
    # same (seed, offset) -> same mask; another offset -> another mask.  (The elementwise pa_relu_dropout_fwd of the exact
    # cuBLAS cross-check path draws its own per-element stream: the epilogue uses one bit-sliced keep word per 32 columns.)
    c3 = torch.empty(M, N, device='cuda')
    ops.gemm_tf32(a.cuda(), w.cuda(), c3, M, N, K, lda=K, ldb=K, ldc=N, bias=bias.cuda(), relu=True, p_drop=p, seed=7, off=3)
    assert torch.equal(c3, c2)
    ops.gemm_tf32(a.cuda(), w.cuda(), c3, M, N, K, lda=K, ldb=K, ldc=N, bias=bias.cuda(), relu=True, p_drop=p, seed=7, off=4)
    assert not torch.equal(c3 > 0, keep)
    both = ((c3 > 0) & keep)[pos].float().mean().item()
    assert abs(both - (1 - p) ** 2) < 1e-2          # independent streams


@pytest.mark.parametrize('M,N,K', [(1196, 512, 1536), (4096, 256, 1024), (256, 1024, 512)])
def test_gemm_dx_form(M, N, K):
    """dx = dy @ W with W [K(contract), N(out)] as stored: A K-major, B MN-major."""
    from plankassembly_b200 import ops
    g = torch.Generator().manual_seed(2)
    dy = torch.randn(M, K, generator=g)
    w = torch.randn(K, N, generator=g) / K ** 0.5
    c = torch.full((M, N), float('nan'), device='cuda')
    ops.gemm_tf32(dy.cuda(), w.cuda(), c, M, N, K, lda=K, ldb=N, ldc=N, b_mn=True)
    check(c, dy.double(), w.double().T, tf32_trunc(dy), tf32_trunc(w).T, f'dX {M}x{N}x{K}')


@pytest.mark.parametrize('Mtok,N,K,split', [(1196, 512, 256, 1), (4096, 1536, 512, 3), (2048, 514, 512, 2)])
def test_gemm_dw_form(Mtok, N, K, split):
    """dW[N,K] = dy^T x: both operands MN-major (stored [tokens, features]), split-K + RED."""
    from plankassembly_b200 import ops
    g = torch.Generator().manual_seed(3)
    dy = torch.randn(Mtok, N, generator=g)
    x = torch.randn(Mtok, K, generator=g)
    c = torch.zeros(N, K, device='cuda')
    ldn = (N + 3) // 4 * 4                   # TMA wants a 16-byte pitch: pad the row, keep the logical N
    dyp = torch.nn.functional.pad(dy, (0, ldn - N)).cuda()
    ops.gemm_tf32(dyp, x.cuda(), c, N, K, Mtok, lda=ldn, ldb=K, ldc=K, a_mn=True, b_mn=True, split_k=split, accumulate=True)
    check(c, dy.double().T, x.double().T, tf32_trunc(dy).T, tf32_trunc(x).T, f'dW {N}x{K}x{Mtok}')


def test_gemm_batched_pointer_scores():
    from plankassembly_b200 import ops
    g = torch.Generator().manual_seed(4)
    for B, T, d in [(3, 256, 512), (4, 64, 128), (2, 128, 512)]:
        pf, h = torch.randn(B, T, d, generator=g), torch.randn(B, T, d, generator=g)
        c = torch.full((B, T, T), float('nan'), device='cuda')
        ops.gemm_tf32(pf.cuda(), h.cuda(), c, T, T, d, lda=d, ldb=d, ldc=T, batch=B, a_batch_rows=T, b_batch_rows=T, c_batch_stride=T * T)
        ref = pf.double() @ h.double().transpose(1, 2)
        assert rel_err(c.cpu(), ref) < 2e-3, (B, T, d)


def test_linear_autograd_matches_torch():
    from plankassembly_b200 import ops
    g = torch.Generator().manual_seed(5)
    B, L, K, N = 4, 299, 512, 1536
    x = torch.randn(B, L, K, generator=g, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(N, K, generator=g, dtype=torch.float64) / K ** 0.5).requires_grad_(True)
    b = torch.randn(N, generator=g, dtype=torch.float64, requires_grad=True)
    ref = torch.nn.functional.linear(x, w, b)
    gy = torch.randn(B, L, N, generator=g, dtype=torch.float64)
    (ref * gy).sum().backward()
    cx, cw, cb = (t.detach().float().cuda().requires_grad_(True) for t in (x, w, b))
    out = ops.Linear.apply(cx, cw, cb, cw.detach(), False, 0.0, False, False)
    assert rel_err(out.cpu(), ref.detach()) < 2e-3
    (out * gy.float().cuda()).sum().backward()
    assert rel_err(cx.grad.cpu(), x.grad) < 2e-3
    assert rel_err(cw.grad.cpu(), w.grad) < 2e-3
    assert rel_err(cb.grad.cpu(), b.grad) < 1e-5


def test_linear_vocab_head_shape():
    """N = 514 (not a multiple of 4): scalar-store epilogue + padded dy in backward."""
    from plankassembly_b200 import ops
    g = torch.Generator().manual_seed(6)
    M, K, N = 1024, 512, 514
    x = torch.randn(M, K, generator=g, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(N, K, generator=g, dtype=torch.float64) / K ** 0.5).requires_grad_(True)
    b = torch.randn(N, generator=g, dtype=torch.float64, requires_grad=True)
    ref = torch.nn.functional.linear(x, w, b)
    gy = torch.randn(M, N, generator=g, dtype=torch.float64)
    (ref * gy).sum().backward()
    cx, cw, cb = (t.detach().float().cuda().requires_grad_(True) for t in (x, w, b))
    out = ops.Linear.apply(cx, cw, cb, cw.detach(), False, 0.0, False, False)
    assert rel_err(out.cpu(), ref.detach()) < 2e-3
    (out * gy.float().cuda()).sum().backward()
    assert rel_err(cx.grad.cpu(), x.grad) < 2e-3
    assert rel_err(cw.grad.cpu(), w.grad) < 2e-3


def test_rounded_operands_remove_truncation_bias():
    """Operands rounded to nearest TF32 by the producer (pa_round_tf32 / epilogue flags) give an
    unbiased product: the error vs fp64 drops well below the ~7.5e-4 of raw (truncated) operands."""
    from plankassembly_b200 import ops
    from plankassembly_b200._lib import call
    g = torch.Generator().manual_seed(7)
    M, N, K = 2048, 512, 512
    a, w = torch.randn(M, K, generator=g).cuda(), (torch.randn(N, K, generator=g) / K ** 0.5).cuda()
    ar, wr = torch.empty_like(a), torch.empty_like(w)
    st = torch.cuda.current_stream().cuda_stream
    call('pa_round_tf32', a.data_ptr(), ar.data_ptr(), a.numel(), st)
    call('pa_round_tf32', w.data_ptr(), wr.data_ptr(), w.numel(), st)
    assert torch.equal(ar.view(torch.int32) & 0x1FFF, torch.zeros_like(ar, dtype=torch.int32))
    assert (ar - a).abs().max() <= a.abs().max() * 2.0 ** -11
    c_raw, c_rn = torch.empty(M, N, device='cuda'), torch.empty(M, N, device='cuda')
    ops.gemm_tf32(a, w, c_raw, M, N, K, lda=K, ldb=K, ldc=N)
    ops.gemm_tf32(ar, wr, c_rn, M, N, K, lda=K, ldb=K, ldc=N)
    ref = a.double().cpu() @ w.double().cpu().T
    e_raw, e_rn = rel_err(c_raw.cpu(), ref), rel_err(c_rn.cpu(), ref)
    print(f'raw (truncated) operands: {e_raw:.2e}; round-to-nearest operands: {e_rn:.2e}')
    assert e_rn < 0.5 * e_raw and e_rn < 4e-4


def test_linear_relu_dropout_bias_gradient_fused():
    """relu+dropout epilogue: backward's fused column-sum kernel must equal the masked gradient's column sums."""
    from plankassembly_b200 import ops
    g = torch.Generator().manual_seed(8)
    M, K, N = 1000, 512, 1024
    x = torch.randn(M, K, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().requires_grad_(True)
    b = torch.randn(N, generator=g).cuda().requires_grad_(True)
    out = ops.Linear.apply(x, w, b, w.detach(), True, 0.2, False, False)
    gy = torch.randn(M, N, generator=g).cuda()
    (out * gy).sum().backward()
    masked = gy * (out.detach() > 0) / 0.8
    assert rel_err(b.grad.cpu(), masked.sum(0).cpu()) < 2e-3


def test_pointer_scores_autograd():
    """Batched pointer scoring (ref models.py:149) forward + both backward GEMMs against fp64 bmm."""
    from plankassembly_b200 import ops
    g = torch.Generator().manual_seed(9)
    for B, T, d in [(3, 256, 512), (4, 64, 128), (2, 128, 256)]:
        pf = torch.randn(B, T, d, generator=g, dtype=torch.float64, requires_grad=True)
        h = torch.randn(B, T, d, generator=g, dtype=torch.float64, requires_grad=True)
        ref = torch.bmm(pf, h.transpose(1, 2))
        w = torch.randn(B, T, T, generator=g, dtype=torch.float64)
        (ref * w).sum().backward()
        cpf, ch = (t.detach().float().cuda().requires_grad_(True) for t in (pf, h))
        out = ops.pointer_scores(cpf, ch, True)
        assert rel_err(out.cpu(), ref.detach()) < 2e-3
        (out * w.float().cuda()).sum().backward()
        assert rel_err(cpf.grad.cpu(), pf.grad) < 2e-3, (B, T, d)
        assert rel_err(ch.grad.cpu(), h.grad) < 2e-3, (B, T, d)


@pytest.mark.parametrize('M,N,K', [(1196, 1536, 512), (64, 514, 512), (4096, 512, 1024), (1000, 1024, 512)])
def test_linear_x3_is_fp32_class(M, N, K):
    """3xTF32 (csrc/split3.cu + one pa_gemm_tf32 over K' = 3K): the exact-mode inference projections.  Error against fp64
    is within an order of magnitude of an fp32 GEMM's and two orders below a plain TF32 GEMM's (~5e-4)."""
    from plankassembly_b200 import ops
    g = torch.Generator().manual_seed(5)
    x = torch.randn(M, K, generator=g)
    w = torch.nn.Parameter(torch.randn(N, K, generator=g) / K ** 0.5)
    b = torch.randn(N, generator=g)
    ops.begin_step()
    with torch.no_grad():
        y = ops.linear_x3(x.cuda(), w.cuda(), b.cuda(), relu=False)
        y_rows = ops.linear_x3(x.cuda(), w.cuda(), b.cuda(), rows=slice(N // 2, N))
        ref32 = torch.nn.functional.linear(x.cuda(), w.detach().cuda(), b.cuda())
    ref = x.double() @ w.detach().double().T + b.double()
    e, e32 = rel_err(y.cpu(), ref), rel_err(ref32.cpu(), ref)
    print(f'x3 {M}x{N}x{K}: rel err vs fp64 {e:.2e} (torch fp32 matmul: {e32:.2e})')
    assert e < 1e-5 and e < 10 * e32 + 1e-7        # measured 5e-6 (fp32 matmul 8e-7, plain TF32 5e-4): tensor-core accumulation order
    assert torch.equal(y_rows, y[:, N // 2:])


@pytest.mark.parametrize('M,d,ff,p', [(1196, 512, 1024, 0.2), (300, 128, 256, 0.0), (4096, 256, 1024, 0.1)])
def test_fused_ffn_matches_the_two_linears(M, d, ff, p):
    """ops.FFN (activation backward fused into the epilogues: relu/dropout bit plane out of the FFN1 GEMM, masked dX + bias
    column sums in linear2's dX GEMM) against the unfused composition with the same RNG state."""
    from plankassembly_b200 import ops
    g = torch.Generator().manual_seed(7)
    x = ops_round(torch.randn(M, d, generator=g)).cuda()
    W1 = (torch.randn(ff, d, generator=g) / d ** 0.5).cuda()
    b1 = torch.randn(ff, generator=g).cuda()
    W2 = (torch.randn(d, ff, generator=g) / ff ** 0.5).cuda()
    w = torch.randn(M, d, generator=g).cuda()
    res = []
    for fused in (True, False):
        xs, p1, pb, p2 = [t.clone().requires_grad_(True) for t in (x, W1, b1, W2)]
        ops.begin_step()
        torch.manual_seed(123)
        if fused:
            out = ops.FFN.apply(xs, p1, pb, ops.tf32_weight(p1), p2, ops.tf32_weight(p2), p)
        else:
            h = ops.linear(xs, p1, pb, relu=True, p_drop=p, tf32=True, round_out=True)
            # round_dx=False: the activation backward of the first linear rounds AFTER scaling by 1/(1-p), exactly once, as
            # the fused epilogue does (rounding before and after the scaling differs by up to one TF32 ulp per element)
            out = ops.linear(h, p2, None, tf32=True, round_dx=False)
        (out * w).sum().backward()
        res.append((out.detach(), xs.grad, p1.grad, pb.grad, p2.grad))
    assert torch.equal(res[0][0], res[1][0])
    for a, b, name in zip(res[0][1:], res[1][1:], ('dx', 'dW1', 'db1', 'dW2')):
        assert rel_err(a.cpu(), b.cpu()) < 2e-5, name


def ops_round(t):
    return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)


def test_shared_input_gradient_accumulates_in_the_epilogue():
    """ops.DxAccum: several Linear nodes reading the SAME activation sum their input gradients inside the dX GEMM epilogues
    (first node stores, the others TMA reduce-add) -- same result as autograd's own accumulation, also on a second backward."""
    from plankassembly_b200 import ops
    g = torch.Generator().manual_seed(11)
    M, K, N = 1196, 512, 1024
    x0 = ops_round(torch.randn(M, K, generator=g)).cuda()
    Ws = [(torch.randn(N, K, generator=g) / K ** 0.5).cuda() for _ in range(3)]
    ws = [torch.randn(M, N, generator=g).cuda() for _ in range(3)]
    res = []
    for shared in (True, False):
        x = x0.clone().requires_grad_(True)
        params = [w.clone().requires_grad_(True) for w in Ws]
        ops.begin_step()
        acc = ops.DxAccum() if shared else None
        outs = [ops.linear(x, p, None, tf32=True, dx_accum=acc) for p in params]
        loss = sum((o * w).sum() for o, w in zip(outs, ws))
        loss.backward(retain_graph=shared)
        g1, gp = x.grad.clone(), [p.grad.clone() for p in params]
        if shared:                      # a second pass over the same graph must start from a fresh buffer
            x.grad = None
            loss.backward()
            assert rel_err(x.grad.cpu(), g1.cpu()) < 1e-6
        res.append((g1, gp))
    assert rel_err(res[0][0].cpu(), res[1][0].cpu()) < 2e-6          # fp32 sums in another order
    for a, b in zip(res[0][1], res[1][1]):
        assert rel_err(a.cpu(), b.cpu()) < 1e-5     # split-K reduce-add order
