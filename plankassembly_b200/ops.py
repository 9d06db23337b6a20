"""torch.autograd wrappers over the C-ABI kernels (include/plank_b200.h).

PyTorch is plumbing here: it owns device memory, streams and the autograd tape; every op below
runs a hand-written sm_100a kernel through ctypes.  All tensors are fp32 CUDA tensors; ids stay
int64 as the reference's batches hold them.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib
import os

from ._lib import AttnBwdArgs, AttnFwdArgs, GemmArgs, call

ATTN_IMPL = {'simt': 0, 'tc': 1}
BWD_TC = os.environ.get('PLANK_B200_ATTN_BWD', 'tc') == 'tc'   # tensor-core backward when the forward was tensor-core


class _ZeroPool:
    """Zero-initialised gradient buffers of one backward pass carved out of ONE pre-zeroed allocation (one memset launch per
    step instead of ~200 small fill kernels).  The demand of a step is learned from the previous one; anything beyond it
    falls back to torch.zeros.  The returned tensors are ordinary views: they keep the block alive for as long as they live."""

    def __init__(self):
        self.buf, self.off, self.demand, self.high = None, 0, 0, 0

    def begin_step(self, device):
        self.high = max(self.high, self.demand)
        self.demand, self.off = 0, 0
        self.buf = torch.zeros(self.high, device=device, dtype=torch.float32) if self.high else None

    def zeros(self, shape, device):
        n = 1
        for x in shape:
            n *= int(x)
        n_al = (n + 63) // 64 * 64                      # 256-byte granules: TMA and float4 alignment
        self.demand += n_al
        b = self.buf
        if b is None or b.device != device or self.off + n_al > b.numel():
            return torch.zeros(tuple(shape), device=device, dtype=torch.float32)
        out = b[self.off:self.off + n].view(tuple(shape))
        self.off += n_al
        return out


zero_pool = _ZeroPool()


def _zeros(shape, device):
    return zero_pool.zeros(shape, device)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.PlankB200Error('plankassembly_b200 ops need CUDA tensors (no CPU fallback); '
                                      'move the model and batch to a B200 device')


class DropoutState:
    """(seed, offset) for every dropout site, drawn from torch's own CUDA generator state.

    Every site of every forward takes the generator's current Philox offset as its stream id and advances the generator, as
    a torch dropout kernel would; backward kernels regenerate the masks from the (seed, offset) they saved.  The masks are
    therefore a function of torch's RNG state: `torch.manual_seed` / `seed_everything` reseed them, and
    `torch.cuda.get_rng_state` / `set_rng_state` (checkpoint resume) reproduce them -- round 1 kept a private counter seeded
    once from `torch.initial_seed()`, which a later `manual_seed` or a resume could not reach."""

    STEP = 4          # torch keeps Philox offsets at multiples of 4

    def next(self):
        gen = torch.cuda.default_generators[torch.cuda.current_device()]
        off = gen.get_offset()
        gen.set_offset(off + self.STEP)
        return gen.initial_seed() & 0xFFFFFFFFFFFFFFFF, off // self.STEP + 1


RNG = DropoutState()


def _ptr_array(ptrs):
    return (C.c_void_p * len(ptrs))(*ptrs)


class EmbedInput(Function):
    """K1 (ref models.py:103-112)."""

    @staticmethod
    def forward(ctx, n_tables, want_r, *args):
        ids, tables = args[:n_tables], args[n_tables:]
        _require_cuda(*ids, *tables)
        ids = [i.contiguous() for i in ids]
        B, S = ids[0].shape
        d = tables[0].shape[1]
        out = torch.empty(B, S, d, device=tables[0].device, dtype=torch.float32)
        out_r = torch.empty_like(out) if want_r else None
        call('pa_embed_input_fwd', _ptr_array([i.data_ptr() for i in ids]), _ptr_array([t.data_ptr() for t in tables]),
             n_tables, B * S, d, out.data_ptr(), _ptr(out_r), _stream())
        ctx.ids = ids
        ctx.shapes = [t.shape for t in tables]
        ctx.n = n_tables
        return (out, out_r) if want_r else out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout, dout_r=None):
        dout = dout.contiguous() if dout_r is None else dout + dout_r
        grads = [_zeros(s, dout.device) for s in ctx.shapes]
        rows = (C.c_int * ctx.n)(*[s[0] for s in ctx.shapes])
        B, S = ctx.ids[0].shape
        call('pa_embed_input_bwd', dout.data_ptr(), _ptr_array([i.data_ptr() for i in ctx.ids]),
             _ptr_array([g.data_ptr() for g in grads]), rows, ctx.n, B * S, dout.shape[-1], _stream())
        return (None, None, *([None] * ctx.n), *grads)


class EmbedOutput(Function):
    """K2 (ref models.py:114-138): value[:, :T-1] shifted right by one behind a zero row."""

    @staticmethod
    def forward(ctx, value, T, dof, want_r, e_val, e_coord, e_pos):
        _require_cuda(value, e_val)
        assert value.stride(1) == 1
        B, d = value.shape[0], e_val.shape[1]
        out = torch.empty(B, T, d, device=e_val.device, dtype=torch.float32)
        out_r = torch.empty_like(out) if want_r else None
        call('pa_embed_output_fwd', value.data_ptr(), value.stride(0), B, T, dof, e_val.data_ptr(), e_coord.data_ptr(),
             e_pos.data_ptr(), d, out.data_ptr(), _ptr(out_r), _stream())
        ctx.value, ctx.T, ctx.dof = value, T, dof
        ctx.shapes = (e_val.shape, e_coord.shape, e_pos.shape)
        return (out, out_r) if want_r else out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout, dout_r=None):
        dout = dout.contiguous() if dout_r is None else dout + dout_r
        gv, gc, gp = (_zeros(s, dout.device) for s in ctx.shapes)
        call('pa_embed_output_bwd', dout.data_ptr(), ctx.value.data_ptr(), ctx.value.stride(0), dout.shape[0], ctx.T,
             ctx.dof, gv.data_ptr(), gc.data_ptr(), gp.data_ptr(), dout.shape[-1], _stream())
        return None, None, None, None, gv, gc, gp


LN_SAVE_PRENORM = os.environ.get('PLANK_B200_LN_SAVE', 'y') == 's'


class AddLayerNorm(Function):
    """y = LayerNorm_eps(x + dropout_p(a + a_bias)); a may be None (final norms); a_bias (may be None) is the
    bias of the linear that produced a, folded in here so that its gradient falls out of the backward kernel.
    want_r: also return the TF32-rounded copy of y that feeds the next tensor-core GEMM;
    round_da: the gradient wrt a feeds tensor-core GEMMs only, so it is written rounded.
    Nothing but the OUTPUT is saved for the backward (the TF32 copy where there is one -- the next GEMM keeps it alive
    anyway): x_hat = (y - beta) / gamma, so the forward writes no pre-norm tensor."""

    @staticmethod
    def forward(ctx, x, a, a_bias, gamma, beta, eps, p_drop, want_r=False, round_da=False):
        _require_cuda(x, gamma)
        x = x.contiguous()
        a = a.contiguous() if a is not None else None
        d = x.shape[-1]
        rows = x.numel() // d
        y = torch.empty_like(x)
        y_r = torch.empty_like(x) if want_r else None
        need_grad = any(ctx.needs_input_grad)
        stats = torch.empty(rows, 2, device=x.device, dtype=torch.float32) if need_grad else None
        s = torch.empty_like(x) if (need_grad and LN_SAVE_PRENORM) else None        # A/B switch: round 1 saved the pre-norm sum
        seed, off = RNG.next() if (p_drop > 0 and a is not None) else (0, 0)
        call('pa_add_ln_fwd', x.data_ptr(), _ptr(a), _ptr(a_bias), gamma.data_ptr(), beta.data_ptr(), eps, p_drop if a is not None else 0.0,
             seed, off, rows, d, y.data_ptr(), _ptr(y_r), _ptr(s), _ptr(stats), _stream())
        ctx.from_y = s is None
        ctx.save_for_backward(s if s is not None else (y_r if y_r is not None else y), stats, gamma, beta)
        ctx.has_a, ctx.p, ctx.seed, ctx.off = a is not None, (p_drop if a is not None else 0.0), seed, off
        ctx.round_da, ctx.has_bias = round_da, a_bias is not None
        return (y, y_r) if want_r else y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, dy_r=None):
        yo, stats, gamma, beta = ctx.saved_tensors
        dy = dy.contiguous()
        dy_r = dy_r.contiguous() if dy_r is not None else None
        d = yo.shape[-1]
        rows = yo.numel() // d
        dx = torch.empty_like(yo)
        da = torch.empty_like(yo) if (ctx.has_a and (ctx.p > 0 or ctx.round_da or ctx.has_bias)) else None
        dgamma = _zeros(gamma.shape, gamma.device)
        dbeta = _zeros(gamma.shape, gamma.device)
        dabias = _zeros(gamma.shape, gamma.device) if ctx.has_bias else None
        call('pa_add_ln_bwd', dy.data_ptr(), _ptr(dy_r), yo.data_ptr(), stats.data_ptr(), gamma.data_ptr(), beta.data_ptr() if ctx.from_y else None, ctx.p,
             ctx.seed, ctx.off, rows, d, dx.data_ptr(), _ptr(da), int(ctx.round_da), dgamma.data_ptr(), dbeta.data_ptr(), _ptr(dabias),
             None, _stream())
        if ctx.has_a and da is None:
            da = dx                      # no dropout, no rounding: same gradient flows to both summands
        return dx, (da if ctx.has_a else None), dabias, dgamma, dbeta, None, None, None, None


class ReluDropout(Function):
    """z <- dropout_p(relu(z)) in place (FFN activation, torch transformer.py _ff_block)."""

    @staticmethod
    def forward(ctx, z, p_drop):
        _require_cuda(z)
        assert z.is_contiguous()
        seed, off = RNG.next() if p_drop > 0 else (0, 0)
        call('pa_relu_dropout_fwd', z.data_ptr(), z.numel(), p_drop, seed, off, _stream())
        ctx.mark_dirty(z)
        ctx.save_for_backward(z)
        ctx.p = p_drop
        return z

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        (out,) = ctx.saved_tensors
        g = g.contiguous().clone()
        call('pa_relu_dropout_bwd', out.data_ptr(), g.data_ptr(), g.numel(), ctx.p, 0, _stream())
        return g, None


def _drop_masks(B, H, Lq, Lk, p_drop, seed, off, device):
    """Bit planes of the attention-dropout keep mask for the tensor-core kernels (dropmask.cu)."""
    lib = _lib.load()
    rows = torch.empty(lib.pa_dropout_mask_words(B * H, Lq, Lk, 0), device=device, dtype=torch.int32)
    cols = torch.empty(lib.pa_dropout_mask_words(B * H, Lq, Lk, 1), device=device, dtype=torch.int32)
    call('pa_dropout_mask', rows.data_ptr(), cols.data_ptr(), B * H, Lq, Lk, p_drop, seed, off, _stream())
    return rows, cols


class _MaskPrefetch:
    """Attention-dropout bit planes of a WHOLE train step generated up front on a side stream.

    The planes depend on nothing but (seed, offset, shape); `dropout_mask_kernel` is pure ALU work (Philox) with no DRAM
    reads, 1.3 ms per step when it runs alone in front of each attention call.  PlankModel.train_step announces the
    attention calls of the step (`plan`), all planes are launched at once on a second stream and run in the shadow of the
    tensor-core / HBM-bound kernels of the main stream; each attention call then `take`s its planes (waiting on their event).

    MEASURED (round 2, profiles/README.md): no gain and a noisier step (22.6-27.7 ms against a steady 22.65 ms) -- every main-
    stream kernel is a persistent one-CTA-per-SM design, so the Philox CTAs do not find idle issue slots, they take them.
    Kept as an opt-in experiment (PLANK_B200_MASK_PREFETCH=1), off by default."""

    def __init__(self):
        self.stream, self.q = None, []

    def plan(self, shapes, p_drop, device):
        self.q = []
        if os.environ.get('PLANK_B200_MASK_PREFETCH', '0') != '1':
            return
        if self.stream is None or self.stream.device != device:
            self.stream = torch.cuda.Stream(device=device)
        main = torch.cuda.current_stream(device)
        with torch.cuda.stream(self.stream):
            for (B, H, Lq, Lk) in shapes:
                seed, off = RNG.next()
                rows, cols = _drop_masks(B, H, Lq, Lk, p_drop, seed, off, device)
                ev = torch.cuda.Event()
                ev.record(self.stream)
                rows.record_stream(main)
                cols.record_stream(main)
                self.q.append(((B, H, Lq, Lk, float(p_drop)), seed, off, (rows, cols), ev))

    def take(self, B, H, Lq, Lk, p_drop):
        if self.q and self.q[0][0] == (B, H, Lq, Lk, float(p_drop)):
            _, seed, off, masks, ev = self.q.pop(0)
            torch.cuda.current_stream().wait_event(ev)
            return seed, off, masks
        self.q = []                     # out of step with the plan: the remaining calls generate their planes in line
        return None


MASKS = _MaskPrefetch()

_KV_LEN_CACHE = []      # [(kpm tensor, version, kv_len)]: one mask serves all layers of a step


def _kv_len(kpm):
    """[B] int32: 1 + index of the last non-PAD key (>= 1).  Batches padded to a fixed length (LineDataset layout) let the
    tensor-core attention kernels skip whole key tiles beyond it; nothing changes numerically (those keys carry -inf)."""
    if kpm is None:
        return None
    try:
        version = kpm._version
    except RuntimeError:                 # inference tensors do not track versions: no caching
        version = None
    if version is not None:
        for t, ver, out in _KV_LEN_CACHE:
            if t is kpm and ver == version:
                return out
    Lk = kpm.shape[1]
    idx = torch.arange(1, Lk + 1, device=kpm.device, dtype=torch.int32)
    out = ((kpm == 0).to(torch.int32) * idx).amax(1).clamp_(min=1).contiguous()
    if version is not None:
        _KV_LEN_CACHE.append((kpm, version, out))
        del _KV_LEN_CACHE[:-4]
    return out


def _attn_fwd(q, k, v, ldq, ldk, ldv, B, H, Lq, Lk, dh, kpm, causal, p_drop, seed, off, impl, want_lse, device, rnd=False, masks=None):
    o = torch.empty(B, Lq, H * dh, device=device, dtype=torch.float32)
    lse = torch.empty(B, H, Lq, device=device, dtype=torch.float32) if want_lse else None
    if masks is None:
        masks = _drop_masks(B, H, Lq, Lk, p_drop, seed, off, device) if (impl == 1 and p_drop > 0) else (None, None)
    a = AttnFwdArgs(q, k, v, ldq, ldk, ldv, o.data_ptr(), H * dh, _ptr(lse), _ptr(kpm), B, H, Lq, Lk, dh, int(causal),
                    dh ** -0.5, p_drop, seed, off, impl, int(rnd), _ptr(masks[0]), _ptr(masks[1]),
                    _ptr(_kv_len(kpm)))
    call('pa_attn_fwd', C.byref(a), _stream())
    return o, lse, masks


def _attn_bwd(q, k, v, ldq, ldk, ldv, o, do, lse, dq, dk, dv, lddq, lddk, lddv, B, H, Lq, Lk, dh, kpm, causal, p_drop,
              seed, off, impl, rnd=False, masks=(None, None), dbias=None):
    delta = torch.empty(B, H, Lq, device=o.device, dtype=torch.float32)
    if impl == 1 and p_drop > 0 and masks[0] is None:
        masks = _drop_masks(B, H, Lq, Lk, p_drop, seed, off, o.device)
    a = AttnBwdArgs(q, k, v, ldq, ldk, ldv, o.data_ptr(), do.data_ptr(), H * dh, lse.data_ptr(), delta.data_ptr(),
                    dq, dk, dv, lddq, lddk, lddv, _ptr(kpm), B, H, Lq, Lk, dh, int(causal), dh ** -0.5, p_drop, seed, off, impl, int(rnd),
                    _ptr(masks[0]), _ptr(masks[1]), _ptr(dbias) if impl == 1 else None,
                    _ptr(_kv_len(kpm)) if impl == 1 else None)
    call('pa_attn_bwd', C.byref(a), _stream(), launches=3)


TC_MAX_LQ, TC_MAX_LK = 1280, 2048      # shared-memory tables of the tensor-core attention kernels (include/plank_b200.h)


def _fit_impl(impl, Lq, Lk):
    """The tensor-core attention kernels keep per-item tables in shared memory (key bias: Lk <= 2048; query statistics of the
    dK/dV kernel: Lq <= 1280).  Longer sequences than any reference config has (MAX_INPUT_LENGTH 1200) are served by the
    fp32 CUDA-core kernels instead of failing."""
    return 0 if (impl == 1 and (Lq > TC_MAX_LQ or Lk > TC_MAX_LK)) else impl


class SelfAttention(Function):
    """K3/K4: attention core over the packed in-projection output qkv [B,L,3d].
    `bias` (may be None) is the in-projection bias PARAMETER: it is not used in the forward (the GEMM epilogue
    already added it) but its gradient -- the column sums of dqkv -- is produced by the backward kernels, so
    the projection's own backward skips it."""

    @staticmethod
    def forward(ctx, qkv, bias, kpm, H, causal, p_drop, impl, rnd=False):
        _require_cuda(qkv)
        qkv = qkv.contiguous()
        B, L, d3 = qkv.shape
        d = d3 // 3
        impl = _fit_impl(impl, L, L)
        pre = MASKS.take(B, H, L, L, p_drop) if (impl == 1 and p_drop > 0) else None
        seed, off, pm = pre if pre is not None else ((*RNG.next(), None) if p_drop > 0 else (0, 0, None))
        base = qkv.data_ptr()
        o, lse, masks = _attn_fwd(base, base + 4 * d, base + 8 * d, d3, d3, d3, B, H, L, L, d // H, kpm, causal, p_drop, seed, off,
                                  impl, any(ctx.needs_input_grad), qkv.device, rnd, pm)
        ctx.save_for_backward(qkv, o, lse, kpm, *masks)
        ctx.cfg = (H, causal, p_drop, seed, off, impl, rnd, bias is not None)
        return o

    @staticmethod
    @once_differentiable
    def backward(ctx, do):
        qkv, o, lse, kpm, m_rows, m_cols = ctx.saved_tensors
        H, causal, p_drop, seed, off, impl, rnd, has_bias = ctx.cfg
        B, L, d3 = qkv.shape
        d = d3 // 3
        dqkv = torch.empty_like(qkv)
        bimpl = impl if BWD_TC else 0
        dbias = _zeros((d3,), qkv.device) if (has_bias and bimpl == 1) else None
        base, g = qkv.data_ptr(), dqkv.data_ptr()
        _attn_bwd(base, base + 4 * d, base + 8 * d, d3, d3, d3, o, do.contiguous(), lse, g, g + 4 * d, g + 8 * d, d3, d3, d3,
                  B, H, L, L, d // H, kpm, causal, p_drop, seed, off, bimpl, rnd, (m_rows, m_cols), dbias)
        if has_bias and dbias is None:
            dbias = dqkv.sum((0, 1))
        return dqkv, dbias, None, None, None, None, None, None


class CrossAttention(Function):
    """K5: queries q [B,Lq,d] against the packed memory projection kv [B,Lk,2d]."""

    @staticmethod
    def forward(ctx, q, kv, bias, kpm, H, p_drop, impl, rnd=False):
        _require_cuda(q, kv)
        q, kv = q.contiguous(), kv.contiguous()
        B, Lq, d = q.shape
        Lk = kv.shape[1]
        impl = _fit_impl(impl, Lq, Lk)
        pre = MASKS.take(B, H, Lq, Lk, p_drop) if (impl == 1 and p_drop > 0) else None
        seed, off, pm = pre if pre is not None else ((*RNG.next(), None) if p_drop > 0 else (0, 0, None))
        kb = kv.data_ptr()
        need = any(ctx.needs_input_grad)
        o, lse, masks = _attn_fwd(q.data_ptr(), kb, kb + 4 * d, d, 2 * d, 2 * d, B, H, Lq, Lk, d // H, kpm, False, p_drop, seed, off,
                                  impl, need, q.device, rnd, pm)
        ctx.save_for_backward(q, kv, o, lse, kpm, *masks)
        ctx.cfg = (H, p_drop, seed, off, impl, rnd, bias is not None)
        return o

    @staticmethod
    @once_differentiable
    def backward(ctx, do):
        q, kv, o, lse, kpm, m_rows, m_cols = ctx.saved_tensors
        H, p_drop, seed, off, impl, rnd, has_bias = ctx.cfg
        B, Lq, d = q.shape
        Lk = kv.shape[1]
        dq, dkv = torch.empty_like(q), torch.empty_like(kv)
        bimpl = impl if BWD_TC else 0
        dbias = _zeros((3 * d,), q.device) if (has_bias and bimpl == 1) else None
        kb, gb = kv.data_ptr(), dkv.data_ptr()
        _attn_bwd(q.data_ptr(), kb, kb + 4 * d, d, 2 * d, 2 * d, o, do.contiguous(), lse, dq.data_ptr(), gb, gb + 4 * d,
                  d, 2 * d, 2 * d, B, H, Lq, Lk, d // H, kpm, False, p_drop, seed, off, bimpl, rnd, (m_rows, m_cols), dbias)
        if has_bias and dbias is None:
            dbias = torch.cat([dq.sum((0, 1)), dkv.sum((0, 1))])
        return dq, dkv, dbias, None, None, None, None, None


def gemm_tf32(a, b, c, M, N, K, *, lda, ldb, ldc, a_mn=False, b_mn=False, bias=None, relu=False, p_drop=0.0, seed=0, off=0,
              alpha=1.0, batch=1, a_batch_rows=0, b_batch_rows=0, c_batch_stride=0, split_k=1, accumulate=False, round_out=False,
              mask_out=None, mask_in=None, colsum=None, mask_scale=0.0):
    """C = alpha * op(A) op(B)^T (+bias)(relu)(dropout) on the tcgen05 tensor cores (TF32 in, FP32 acc).
    mask_out / mask_in / colsum / mask_scale: the fused FFN activation backward (include/plank_b200.h, pa_gemm_args)."""
    g = GemmArgs(a.data_ptr(), lda, int(a_mn), b.data_ptr(), ldb, int(b_mn), c.data_ptr(), ldc, _ptr(bias), int(relu), p_drop,
                 seed, off, alpha, M, N, K, batch, a_batch_rows, b_batch_rows, c_batch_stride, split_k, int(accumulate),
                 int(round_out), _ptr(mask_out), _ptr(mask_in), _ptr(colsum), mask_scale)
    call('pa_gemm_tf32', C.byref(g), _stream())
    return c


def _split_k(tiles, k_blocks):
    """Split the contraction of a weight-gradient GEMM so ~one wave of CTAs covers the 148 SMs."""
    sk = max(1, 148 // max(1, tiles))
    return max(1, min(sk, k_blocks // 8 if k_blocks >= 8 else 1))


# ---- derived weight copies for the tensor cores.  Both caches hang off the parameter through a WeakKeyDictionary (they
# die with it) and are refreshed ONCE PER STEP, whatever happened to the weights in between: `begin_step()` (called by
# PlankModel.train_step / eval_step) opens a new epoch and a copy made in an older epoch is rebuilt on first use.  No
# reliance on Tensor._version, which updates made through `.data` (EMA swaps, hand-written optimizers) do not bump.
import weakref


class _IdWeakCache:
    """parameter -> entry, keyed by IDENTITY (tensors compare elementwise, so they cannot be WeakKeyDictionary keys) and
    dropped by a finalizer when the parameter dies, so that neither the parameter nor its copy is pinned in GPU memory."""

    def __init__(self):
        self.d = {}

    def get(self, t):
        ent = self.d.get(id(t))
        return ent[1] if ent is not None and ent[0]() is t else None

    def set(self, t, value):
        key = id(t)
        old = self.d.get(key)
        if old is None or old[0]() is not t:
            weakref.finalize(t, self._drop, key, weakref.ref(t))
        self.d[key] = (weakref.ref(t), value)

    def _drop(self, key, ref):
        ent = self.d.get(key)
        if ent is not None and ent[0] is ref:
            del self.d[key]

    def __len__(self):
        return len(self.d)


_EPOCH = [0]
_W_TF32 = _IdWeakCache()      # parameter -> [epoch, rounded copy]           (training GEMMs)
_W_X3 = _IdWeakCache()        # parameter -> [epoch, [N, 3K] hi|hi|lo copy]  (exact-mode inference GEMMs)


def begin_step():
    _EPOCH[0] += 1


def mark_shadow_fresh(W, buf):
    """An optimizer that writes the rounded copy itself (optim.FusedAdam) registers it for the NEXT step."""
    _W_TF32.set(W, [_EPOCH[0] + 1, buf])


def tf32_weight(W):
    """TF32-rounded shadow copy of a parameter (rebuilt at most once per step)."""
    ent = _W_TF32.get(W)
    if ent is None or ent[0] < _EPOCH[0] or ent[1].device != W.device or ent[1].shape != W.shape:
        buf = ent[1] if (ent is not None and ent[1].shape == W.shape and ent[1].device == W.device) else torch.empty_like(W)
        call('pa_round_tf32', W.data_ptr(), buf.data_ptr(), W.numel(), _stream())
        ent = [_EPOCH[0], buf]
        _W_TF32.set(W, ent)
    return ent[1]


def split3(x2, weights=False, out=None):
    """[rows, K] -> [rows, 3K] error-compensated TF32 operand (csrc/split3.cu)."""
    rows, K = x2.shape
    if out is None:
        out = torch.empty(rows, 3 * K, device=x2.device, dtype=torch.float32)
    call('pa_split3_tf32', x2.data_ptr(), x2.stride(0), out.data_ptr(), rows, K, int(weights), _stream())
    return out


def x3_weight(W):
    """[N, 3K] = [hi | hi | lo] copy of a weight for the 3xTF32 GEMMs (rebuilt at most once per step; the buffer itself is
    stable, so CUDA graphs that captured its address stay valid)."""
    ent = _W_X3.get(W)
    if ent is None or ent[0] < _EPOCH[0] or ent[1].device != W.device or ent[1].shape[0] != W.shape[0]:
        with torch.inference_mode(False):
            buf = ent[1] if (ent is not None and ent[1].device == W.device and ent[1].shape == (W.shape[0], 3 * W.shape[1])) \
                else torch.empty(W.shape[0], 3 * W.shape[1], device=W.device, dtype=torch.float32)
        Wd = W.detach()
        split3(Wd if Wd.is_contiguous() else Wd.contiguous(), True, buf)
        ent = [_EPOCH[0], buf]
        _W_X3.set(W, ent)
    return ent[1]


def linear_x3(x, W, b, rows=None, relu=False, out=None):
    """y = x W[rows]^T + b[rows] in fp32-class accuracy on the tensor cores (3xTF32, see csrc/split3.cu).  Inference only."""
    K = W.shape[1]
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1 or x2.stride(0) % 4:
        x2 = x2.contiguous()
    W3 = x3_weight(W)
    bv = b
    if rows is not None:
        W3 = W3[rows]
        bv = b[rows] if b is not None else None
    M, N = x2.shape[0], W3.shape[0]
    y = out if out is not None else torch.empty(M, N, device=x.device, dtype=torch.float32)
    gemm_tf32(split3(x2), W3, y, M, N, 3 * K, lda=3 * K, ldb=3 * K, ldc=N, bias=bv.detach() if bv is not None else None, relu=relu)
    return y.view(*x.shape[:-1], N)


class DxAccum:
    """One input-gradient buffer shared by SEVERAL Linear nodes that read the same activation (the encoder memory feeds the
    cross-attention K/V projection of every decoder layer): the first backward to run stores its dx, the others add theirs
    into it in the GEMM epilogue (TMA reduce-add) and hand autograd no gradient of their own, so the sum costs no extra
    kernels (it was five elementwise adds of the [B,S,d] gradient per step).  Stream order makes the buffer complete before
    the producer's backward reads it: autograd runs that node only after all the sharing nodes have been enqueued."""

    def __init__(self):
        self.buf, self.n_fwd, self.n_bwd = None, 0, 0


class Linear(Function):
    """y = x W^T + b (optionally relu + dropout fused in the GEMM epilogue) on the TF32 tensor cores.
    x must already be TF32-rounded by its producer; W_r is the rounded shadow of W (W itself only routes
    the gradient).  Backward: dx = dy W_r (B operand MN-major: W as stored), dW = dy^T x (both operands
    MN-major, split-K over the tokens with fp32 RED), db = column sums of dy.
    round_out / round_dx: y / dx feed tensor-core operands only and are written rounded."""

    @staticmethod
    def forward(ctx, x, W, b, W_r, relu, p_drop, round_out, round_dx, *extra):       # extra: an optional DxAccum
        _require_cuda(x, W)
        dx_accum = extra[0] if extra else None
        # extra[1] = (r0, r1, DxAccum): W is the FULL parameter, W_r its rows r0:r1; the weight gradient of the slice is written
        # into rows r0:r1 of ONE zeroed full-size buffer shared by the nodes that use disjoint row ranges of the parameter (the q
        # and k/v rows of a cross-attention in-projection) -- no slice-backward fills / copies and no add of two full gradients
        ctx.w_rows = extra[1] if len(extra) > 1 else None
        if ctx.w_rows is not None:
            ctx.w_rows[2].n_fwd += 1
        ctx.dx_accum, ctx.n_extra = dx_accum, len(extra)
        if dx_accum is not None:
            dx_accum.n_fwd += 1
        K, N = W_r.shape[1], W_r.shape[0]
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        assert W_r.stride(1) == 1
        M = x2.shape[0]
        y = torch.empty(M, N, device=x.device, dtype=torch.float32)
        seed, off = RNG.next() if p_drop > 0 else (0, 0)
        gemm_tf32(x2, W_r, y, M, N, K, lda=K, ldb=W_r.stride(0), ldc=N, bias=b, relu=relu, p_drop=p_drop, seed=seed, off=off,
                  round_out=round_out)
        ctx.save_for_backward(x2, W_r, y if (relu or p_drop > 0) else None)
        ctx.w_full_rows = W.shape[0]
        ctx.cfg = (relu, p_drop, b is not None, x.shape, round_dx)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x2, W_r, y = ctx.saved_tensors
        relu, p_drop, has_b, xshape, round_dx = ctx.cfg
        N, K = W_r.shape
        M = x2.shape[0]
        dy2 = dy.reshape(M, N)
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        want_db = has_b and ctx.needs_input_grad[2]
        db = None
        if y is not None:
            dy2 = dy2.clone()
            if want_db and N % 4 == 0 and N // 4 <= 256 and 256 % (N // 4) == 0:
                db = _zeros((N,), dy.device)     # bias grad = column sums, fused
                call('pa_relu_dropout_bwd_colsum', y.data_ptr(), dy2.data_ptr(), M, N, p_drop, 1, db.data_ptr(), _stream())
            else:
                call('pa_relu_dropout_bwd', y.data_ptr(), dy2.data_ptr(), dy2.numel(), p_drop, 1, _stream())
        if want_db and db is None:
            db = dy2.sum(0)
        ldn = N
        if N % 4:                         # TMA needs a 16-byte row pitch (vocab head: N = 514)
            ldn = (N + 3) // 4 * 4
            dy2 = torch.nn.functional.pad(dy2, (0, ldn - N))
        dx = dW = None
        acc = ctx.dx_accum
        if ctx.needs_input_grad[0] and acc is not None and not round_dx and K % 4 == 0:
            first = acc.buf is None
            if first:
                acc.buf = torch.empty(M, K, device=dy.device, dtype=torch.float32)
            gemm_tf32(dy2, W_r, acc.buf, M, K, N, lda=ldn, ldb=W_r.stride(0), ldc=K, b_mn=True, accumulate=not first)
            dx = acc.buf.view(xshape) if first else None        # the later nodes add in place: no gradient of their own
            acc.n_bwd += 1
            if acc.n_bwd == acc.n_fwd:                           # a second backward over the same graph starts afresh
                acc.buf, acc.n_bwd = None, 0
        elif ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, device=dy.device, dtype=torch.float32)
            gemm_tf32(dy2, W_r, dx, M, K, N, lda=ldn, ldb=W_r.stride(0), ldc=K, b_mn=True, round_out=round_dx)
            dx = dx.view(xshape)
        if ctx.needs_input_grad[1]:
            dW_full = None
            if ctx.w_rows is not None:
                r0, r1, wacc = ctx.w_rows
                first = wacc.buf is None
                if first:
                    wacc.buf = _zeros((ctx.w_full_rows, K), dy.device)
                dW, dW_full = wacc.buf[r0:r1], (wacc.buf if first else None)
                wacc.n_bwd += 1
                if wacc.n_bwd == wacc.n_fwd:
                    wacc.buf, wacc.n_bwd = None, 0
            else:
                dW = _zeros((N, K), dy.device)
            tiles = ((N + 127) // 128) * ((K + 255) // 256 if K % 256 == 0 else (K + 127) // 128)
            gemm_tf32(dy2, x2, dW, N, K, M, lda=ldn, ldb=K, ldc=K, a_mn=True, b_mn=True,
                      split_k=_split_k(tiles, (M + 31) // 32), accumulate=True)
            if ctx.w_rows is not None:
                dW = dW_full
        return (dx, dW, db, None, None, None, None, None) + (None,) * ctx.n_extra


class FFN(Function):
    """out = W2 . dropout_p(relu(W1 x + b1))  (torch transformer.py `_ff_block`; linear2's bias is folded into the following
    residual+LayerNorm kernel) as ONE autograd node on the TF32 tensor cores, so that the activation's backward never makes
    its own pass over the [tokens, ff] gradient: the FFN1 GEMM epilogue leaves one bit per hidden unit (active and kept),
    and the epilogue of linear2's dX GEMM applies it (x 1/(1-p)), rounds, and accumulates the column sums = db1.
    x must be TF32-rounded; W1_r / W2_r are the rounded shadows of W1 / W2."""

    @staticmethod
    def forward(ctx, x, W1, b1, W1_r, W2, W2_r, p_drop):
        _require_cuda(x, W1, W2)
        K, ff, d_out = W1.shape[1], W1.shape[0], W2.shape[0]
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        M = x2.shape[0]
        h = torch.empty(M, ff, device=x.device, dtype=torch.float32)
        mask = torch.empty(M, ff // 32, device=x.device, dtype=torch.int32)
        seed, off = RNG.next() if p_drop > 0 else (0, 0)
        gemm_tf32(x2, W1_r, h, M, ff, K, lda=K, ldb=W1_r.stride(0), ldc=ff, bias=b1, relu=True, p_drop=p_drop, seed=seed, off=off,
                  round_out=True, mask_out=mask)
        out = torch.empty(M, d_out, device=x.device, dtype=torch.float32)
        gemm_tf32(h, W2_r, out, M, d_out, ff, lda=ff, ldb=W2_r.stride(0), ldc=d_out)
        ctx.save_for_backward(x2, h, mask, W1_r, W2_r)
        ctx.cfg = (p_drop, x.shape)
        return out.view(*x.shape[:-1], d_out)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x2, h, mask, W1_r, W2_r = ctx.saved_tensors
        p_drop, xshape = ctx.cfg
        M, K = x2.shape
        ff, d_out = W1_r.shape[0], W2_r.shape[0]
        dy = dout.reshape(M, d_out)
        if not dy.is_contiguous():
            dy = dy.contiguous()
        dev = dy.device
        # dh = dy W2, masked by the activation plane, rounded (it only feeds tensor-core operands); db1 = its column sums
        dh = torch.empty(M, ff, device=dev, dtype=torch.float32)
        db1 = _zeros((ff,), dev)
        gemm_tf32(dy, W2_r, dh, M, ff, d_out, lda=d_out, ldb=W2_r.stride(0), ldc=ff, b_mn=True, round_out=True,
                  mask_in=mask, mask_scale=1.0 / (1.0 - p_drop) if p_drop > 0 else 1.0, colsum=db1)
        kb = (M + 31) // 32
        dW2 = _zeros((d_out, ff), dev)
        gemm_tf32(dy, h, dW2, d_out, ff, M, lda=d_out, ldb=ff, ldc=ff, a_mn=True, b_mn=True,
                  split_k=_split_k(((d_out + 127) // 128) * ((ff + 255) // 256), kb), accumulate=True)
        dx = torch.empty(M, K, device=dev, dtype=torch.float32)
        gemm_tf32(dh, W1_r, dx, M, K, ff, lda=ff, ldb=W1_r.stride(0), ldc=K, b_mn=True)
        dW1 = _zeros((ff, K), dev)
        gemm_tf32(dh, x2, dW1, ff, K, M, lda=ff, ldb=K, ldc=K, a_mn=True, b_mn=True,
                  split_k=_split_k(((ff + 127) // 128) * ((K + 255) // 256), kb), accumulate=True)
        return dx.view(xshape), dW1, db1, None, dW2, None, None


class PointerScores(Function):
    """Raw pointer scores lp[b] = pf[b] h[b]^T (ref models.py:149; the 1/d and the masking live in the
    distribution kernel) as one batched tcgen05 TF32 GEMM; backward = two batched GEMMs with MN-major operands:
    dpf[b] = dlp[b] h[b],  dh[b] = dlp[b]^T pf[b]."""

    @staticmethod
    def forward(ctx, pf, h):
        _require_cuda(pf, h)
        pf, h = pf.contiguous(), h.contiguous()
        B, T, d = pf.shape
        lp = torch.empty(B, T, T, device=pf.device, dtype=torch.float32)
        gemm_tf32(pf, h, lp, T, T, d, lda=d, ldb=d, ldc=T, batch=B, a_batch_rows=T, b_batch_rows=T, c_batch_stride=T * T)
        ctx.save_for_backward(pf, h)
        return lp

    @staticmethod
    @once_differentiable
    def backward(ctx, dlp):
        pf, h = ctx.saved_tensors
        B, T, d = pf.shape
        dlp = dlp.contiguous()
        dpf, dh = torch.empty_like(pf), torch.empty_like(h)
        gemm_tf32(dlp, h, dpf, T, d, T, lda=T, ldb=d, ldc=d, b_mn=True, batch=B, a_batch_rows=T, b_batch_rows=T,
                  c_batch_stride=T * d, round_out=True)
        gemm_tf32(dlp, pf, dh, T, d, T, lda=T, ldb=d, ldc=d, a_mn=True, b_mn=True, batch=B, a_batch_rows=T, b_batch_rows=T,
                  c_batch_stride=T * d)
        return dpf, dh


def pointer_scores(pf, h, tf32):
    """lp = pf @ h^T per sequence: tensor-core batched GEMM on the TF32 path (needs T % 32 == 0 for the
    MN-major backward operands), fp32 cuBLAS bmm on the exact path."""
    if tf32 and pf.shape[1] % 32 == 0:
        return PointerScores.apply(pf, h)
    return torch.bmm(pf, h.transpose(1, 2))


GEMM_IMPL = os.environ.get('PLANK_B200_GEMM', 'tc')
FFN_FUSED = os.environ.get('PLANK_B200_FFN_FUSED', '1') == '1'        # A/B switch for ops.FFN
DX_ACCUM = os.environ.get('PLANK_B200_DX_ACCUM', '1') == '1'          # A/B switch for ops.DxAccum (shared input gradients)


def linear(x, W, b, rows=None, relu=False, p_drop=0.0, tf32=False, round_out=False, round_dx=False, bias_grad=True, dx_accum=None,
           w_accum=None):
    """Dense projection y = x W[rows]^T + b[rows].
    tf32=True : our tcgen05 TF32 GEMM (training path; x must be a TF32-rounded tensor).
    tf32=False: fp32-class result.  Without autograd (inference) it is the same tcgen05 kernel fed error-compensated
                3xTF32 operands (linear_x3); with autograd (the PLANK_B200_GEMM=cublas exact TRAINING path that the
                parity tests use as a cross-check) cuBLAS fp32 through torch."""
    Wv = W if rows is None else W[rows]
    bv = b if (rows is None or b is None) else b[rows]
    if tf32 and not bias_grad and bv is not None:
        bv = bv.detach()                 # bias is added here, its gradient comes from the consumer kernel
    if tf32:
        W_r = tf32_weight(W)
        if rows is not None and w_accum is not None and DX_ACCUM:
            return Linear.apply(x, W, bv, W_r[rows], relu, p_drop, round_out, round_dx, dx_accum, (rows.start, rows.stop, w_accum))
        return Linear.apply(x, Wv, bv, W_r if rows is None else W_r[rows], relu, p_drop, round_out, round_dx, dx_accum)
    if GEMM_IMPL == 'tc' and not torch.is_grad_enabled() and p_drop == 0.0 and W.shape[1] % 4 == 0:
        return linear_x3(x, W, b, rows, relu)        # inference: our own tensor-core kernel in 3xTF32, no cuBLAS
    y = torch.nn.functional.linear(x, Wv, bv)
    if relu:
        y = ReluDropout.apply(y, p_drop)
    return y


class DistLoss(Function):
    """K9/K10 (ref models.py:156-166, 219-227): loss, accuracy and argmax without the [B,T,V+T] tensor.
    lp holds the RAW pointer scores pf.h^T; the kernel applies 1/d and the 1e-6 fill."""

    @staticmethod
    def forward(ctx, lv, lp, sw, label, pad, inv_d, rnd=False):
        _require_cuda(lv, lp, sw, label)
        lv, lp, sw, label = lv.contiguous(), lp.contiguous(), sw.contiguous(), label.contiguous()
        B, T, V = lv.shape
        rowstat = torch.empty(B * T, 4, device=lv.device, dtype=torch.float32)
        predict = torch.empty(B, T, device=lv.device, dtype=torch.int64)
        accum = torch.zeros(3, device=lv.device, dtype=torch.float32)
        call('pa_dist_loss_fwd', lv.data_ptr(), lp.data_ptr(), sw.data_ptr(), label.data_ptr(), B, T, V, pad, inv_d,
             rowstat.data_ptr(), predict.data_ptr(), accum.data_ptr(), _stream())
        loss = accum[0] / accum[1]
        accuracy = accum[2] / (accum[1] + 1e-10)
        ctx.save_for_backward(lv, lp, sw, label, rowstat, accum)
        ctx.cfg = (pad, inv_d, rnd)
        ctx.mark_non_differentiable(accuracy, predict)
        return loss, accuracy, predict

    @staticmethod
    @once_differentiable
    def backward(ctx, gloss, _gacc, _gpred):
        lv, lp, sw, label, rowstat, accum = ctx.saved_tensors
        pad, inv_d, rnd = ctx.cfg
        B, T, V = lv.shape
        dlv, dlp, dsw = torch.empty_like(lv), torch.empty_like(lp), torch.empty_like(sw)
        gloss = gloss.contiguous().to(torch.float32)
        call('pa_dist_loss_bwd', lv.data_ptr(), lp.data_ptr(), sw.data_ptr(), label.data_ptr(), rowstat.data_ptr(),
             accum.data_ptr(), gloss.data_ptr(), B, T, V, pad, inv_d, dlv.data_ptr(), dlp.data_ptr(), dsw.data_ptr(), int(rnd),
             _stream())
        return dlv, dlp, dsw, None, None, None, None


def dist_train_full(lv, lp, sw, inv_d):
    """Materialise the training distribution [B,T,V+T] (parity tests only)."""
    B, T, V = lv.shape
    out = torch.empty(B, T, V + T, device=lv.device, dtype=torch.float32)
    call('pa_dist_train_full', lv.contiguous().data_ptr(), lp.contiguous().data_ptr(), sw.contiguous().data_ptr(), B, T, V,
         inv_d, out.data_ptr(), _stream())
    return out
