"""KV-cached greedy decoder (inference path of PlankModel.eval_step).

The reference regenerates everything on every step (ref models.py:284-307): it re-embeds the
whole prefix, re-runs all decoder layers over all positions, re-projects the encoder memory to
K/V in every layer and builds the full [B,t,V+t] distribution to read its last row, with >= 3
host syncs per step.  Here the per-sequence state is persistent in HBM:

  cross K/V  [L_dec][B,S,2d]   projected ONCE from the encoder memory,
  self  K/V  [L_dec][B,Tmax,d] appended by the attention kernel at each step,
  Hfin       [B,Tmax,d]        final-normed hidden of every emitted position = pointer keys,
  samples / attach [B,Tmax] int64, first_end [B] int32, step counter int32 (device).

One decode step (embedding, 6 x [self-attn, cross-attn, FFN], heads, sampling) touches one new
position only and is captured ONCE as a CUDA graph; the step index lives in device memory, so the
same graph is replayed for every step with no host work in between.  END detection stays on the
device; the host looks at it every `poll` steps.  The reference keeps decoding rows that already
emitted END until ALL rows have (ref models.py:306) and returns those tokens too; the same happens
here and the result is truncated to the reference's stop step max_b(first_end)+1, so
`samples`/`attach` are identical.  Everything is fp32 (bit-exact tokens are the contract).
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F

import ctypes as C

from . import _lib, ops
from ._lib import GemmArgs, call


def _stream():
    return torch.cuda.current_stream().cuda_stream


class GreedyDecoder:
    def __init__(self, model, poll=16):
        self.m = model
        self.poll = poll
        # 'graph' (default): one CUDA graph of ~100 small kernels per step; 'eager': the same kernels launched one by one;
        # 'fused': the whole loop as ONE persistent cooperative kernel (csrc/decode_fused.cu) -- correct, but not yet
        # faster than the graph at batch 64 (profiles/r1_decode_fused.md)
        self.mode = os.environ.get('PLANK_B200_DECODE', 'graph')
        if os.environ.get('PLANK_B200_DECODE_GRAPH', '1') != '1':
            self.mode = 'eager'
        self.use_graph = self.mode == 'graph'
        # sequences decoded together >= this: the step's projections run on the tensor cores as 3xTF32 GEMMs
        # (csrc/split3.cu + pa_gemm_tf32) instead of the small-M fp32 kernel, which re-stages W for every 32 rows
        self.tc_min_batch = int(os.environ.get('PLANK_B200_DECODE_TC_MIN', '128'))
        # large batches are decoded as independent CHAINS of this many sequences, one CUDA stream (= one parallel branch of
        # the captured graph) each: a chain's small-M GEMMs occupy 6-24 SMs, so several chains together fill the machine
        # and one chain's K/V streaming (HBM) overlaps another chain's projections (tensor cores)
        # Measured (profiles/README.md): B=256 -> 8 chains of 32 with the fp32 FMA projections 207 k tok/s (2 chains of 128
        # with 3xTF32: 112 k); B=1024 -> 8 chains of 128 with 3xTF32 251 k (16 chains of 64 with fp32 FMA: 217 k).
        # Default: chains of 128 from batch 1024 on, chains of 32 below; PLANK_B200_DECODE_CHAIN_ROWS overrides.
        self.chain_rows = int(os.environ.get('PLANK_B200_DECODE_CHAIN_ROWS', '0'))
        self._key = None
        self.graph = None
        self._lin_out = {}
        self._x3 = {}

    # ------------------------------------------------------------------ persistent state
    def _buffers(self, B, S, device):
        m = self.m
        # the captured graph bakes in the address of every parameter it reads: key on all of them
        key = (B, S, str(device), hash(tuple(p.data_ptr() for p in m.parameters())))
        if self._key == key:
            return
        # persistent state is allocated as ORDINARY tensors even when the first call happens under torch.inference_mode
        # (Lightning's test loop): they are reset in place by later calls, which may run outside inference mode
        with torch.inference_mode(False):
            d, T, L = m.num_model, m.max_output_length, len(m.decoder.layers)
            f32 = dict(device=device, dtype=torch.float32)
            self.cross_kv = [torch.empty(B, S, 2 * d, **f32) for _ in range(L)]
            self.self_k = [torch.empty(B, T, d, **f32) for _ in range(L)]
            self.self_v = [torch.empty(B, T, d, **f32) for _ in range(L)]
            self.hfin = torch.empty(B, T, d, **f32)
            self.y = torch.empty(B, d, **f32)
            self.y2 = torch.empty(B, d, **f32)
            self.o = torch.empty(B, d, **f32)
            self.kpm = torch.empty(B, S, device=device, dtype=torch.uint8)
            self.kv_len = torch.empty(B, device=device, dtype=torch.int32)
            self.samples = torch.empty(B, T, device=device, dtype=torch.int64)
            self.attach = torch.empty(B, T, device=device, dtype=torch.int64)
            self.first_end = torch.empty(B, device=device, dtype=torch.int32)
            self.t_dev = torch.zeros(1, device=device, dtype=torch.int32)
            self.state = torch.zeros(16, device=device, dtype=torch.int32)
            ws = _lib.load().pa_decode_fused_workspace(B, d, m.decoder.layers[0].linear1.weight.shape[0], m.vocab_size)
            self.part = torch.empty(ws // 4, device=device, dtype=torch.float32)
        self._key, self.graph = key, None
        self._lin_out, self._x3 = {}, {}
        rows = self.chain_rows if self.chain_rows > 0 else (128 if B >= 1024 else 32)
        n_chains = max(1, min(16, B // rows))
        per = -(-B // n_chains)
        self.chains = [(c0, min(B, c0 + per)) for c0 in range(0, B, per)]
        self.streams = [torch.cuda.Stream(device=device) for _ in self.chains] if len(self.chains) > 1 else []
        self.tc3 = per >= self.tc_min_batch and ops.GEMM_IMPL == 'tc'

    def _fused_args(self, B, S):
        """Argument block of pa_decode_fused (raw device pointers of the canonical fp32 parameters)."""
        m = self.m
        d, T = m.num_model, m.max_output_length
        a = _lib.DecodeFusedArgs()
        a.B, a.S, a.T, a.d, a.H, a.ff, a.V, a.L = (B, S, T, d, m.num_head, m.decoder.layers[0].linear1.weight.shape[0], m.vocab_size,
                                                  len(m.decoder.layers))
        a.dof, a.end_token, a.layer_eps, a.final_eps = m.num_output_dof, m.token.END, float(m.layer_eps), 1e-5
        keep = []                                    # tensors that must outlive the launch

        def p(t):
            t = t.detach()
            if not t.is_contiguous():
                t = t.contiguous()
            keep.append(t)
            return t.data_ptr()

        for li, l in enumerate(m.decoder.layers):
            sa, ca, ly = l.self_attn, l.multihead_attn, a.layers[li]
            ly.w_sqkv, ly.b_sqkv, ly.w_so, ly.b_so = p(sa.in_proj_weight), p(sa.in_proj_bias), p(sa.out_proj.weight), p(sa.out_proj.bias)
            ly.g1, ly.be1, ly.g2, ly.be2, ly.g3, ly.be3 = (p(l.norm1.weight), p(l.norm1.bias), p(l.norm2.weight), p(l.norm2.bias),
                                                           p(l.norm3.weight), p(l.norm3.bias))
            ly.w_cq, ly.b_cq = p(ca.in_proj_weight), p(ca.in_proj_bias)          # rows 0..d-1 of the packed projection
            ly.w_co, ly.b_co = p(ca.out_proj.weight), p(ca.out_proj.bias)
            ly.w_f1, ly.b_f1, ly.w_f2, ly.b_f2 = p(l.linear1.weight), p(l.linear1.bias), p(l.linear2.weight), p(l.linear2.bias)
            ly.self_k, ly.self_v, ly.cross_kv = self.self_k[li].data_ptr(), self.self_v[li].data_ptr(), self.cross_kv[li].data_ptr()
        a.gf, a.bf = p(m.decoder.norm.weight), p(m.decoder.norm.bias)
        # vocab | pointer-feature | switch heads as one [V+d+1, d] projection
        a.w_heads = p(torch.cat([m.vocab_head.weight, m.pointer_head.weight, m.switch_head.weight], 0))
        a.b_heads = p(torch.cat([m.vocab_head.bias, m.pointer_head.bias, m.switch_head.bias], 0))
        a.e_val, a.e_coord, a.e_pos = (p(m.input_embeddings['input_value'].weight), p(m.query_coord_embedding.weight),
                                       p(m.query_pos_embedding.weight))
        a.kpm, a.y, a.o = self.kpm.data_ptr(), self.y.data_ptr(), self.o.data_ptr()
        a.part, a.part_bytes = self.part.data_ptr(), self.part.numel() * 4
        a.hfin, a.samples, a.attach = self.hfin.data_ptr(), self.samples.data_ptr(), self.attach.data_ptr()
        a.first_end, a.state = self.first_end.data_ptr(), self.state.data_ptr()
        a.chains = int(os.environ.get('PLANK_B200_DECODE_CHAINS', '0'))
        a.profile = int(os.environ.get('PLANK_B200_DECODE_PROF', '0'))
        return a, keep

    def _split(self, x, key):
        """[M, K] activation -> persistent [M, 3K] hi|lo|hi operand of the 3xTF32 GEMMs (one per distinct input of a step)."""
        M, K = x.shape
        buf = self._x3.get(key)
        if buf is None or buf.shape != (M, 3 * K) or buf.device != x.device:
            with torch.inference_mode(False):
                buf = self._x3[key] = torch.empty(M, 3 * K, device=x.device, dtype=torch.float32)
        return ops.split3(x, False, buf)

    def _lin(self, x, W, b, key, relu=False, rows=None, x3=None):
        """y = x W[rows]^T + b[rows], fp32-class: pa_gemm_skinny_f32 (fp32 FMA) for small batches, the tcgen05 GEMM on
        3xTF32 operands (x3 = the split of x, shared by the projections that read the same x) for large ones.
        Outputs live in persistent buffers (CUDA-graph capture)."""
        M, K = x.shape
        Wv = W if rows is None else W[rows]
        bv = b if (rows is None or b is None) else b[rows]
        N = Wv.shape[0]
        out = self._lin_out.get(key)
        if out is None or out.shape != (M, N) or out.device != x.device:
            with torch.inference_mode(False):
                out = self._lin_out[key] = torch.empty(M, N, device=x.device, dtype=torch.float32)
        if self.tc3 and N >= 64:
            W3 = ops.x3_weight(W)
            if rows is not None:
                W3 = W3[rows]
            if x3 is None:
                x3 = self._split(x, key)
            g = GemmArgs(x3.data_ptr(), 3 * K, 0, W3.data_ptr(), 3 * K, 0, out.data_ptr(), N, bv.data_ptr() if bv is not None else None,
                         int(relu), 0.0, 0, 0, 1.0, M, N, 3 * K, 1, 0, 0, 0, 1, 0, 0)
            call('pa_gemm_tf32', C.byref(g), _stream())
            return out
        call('pa_gemm_skinny_f32', x.data_ptr(), x.stride(0), Wv.data_ptr(), Wv.stride(0), bv.data_ptr() if bv is not None else None,
             out.data_ptr(), N, M, N, K, int(relu), _stream())
        return out

    def _add_ln(self, x, a, norm, eps, out):
        call('pa_add_ln_fwd', x.data_ptr(), a.data_ptr() if a is not None else None, None, norm.weight.data_ptr(),
             norm.bias.data_ptr(), eps, 0.0, 0, 0, x.shape[0], x.shape[1], out.data_ptr(), None, None, None, _stream())
        return out

    # ------------------------------------------------------------------ one step
    def _step(self, t, t_dev):
        """Issue the kernels of decode step t.  t_dev: device pointer of the step counter (graph mode) or None.
        Every chain of sequences runs on its own stream (forked from / joined into the current one)."""
        if len(self.chains) == 1:
            self._step_rows(0, *self.chains[0], t, t_dev)
        else:
            main = torch.cuda.current_stream()
            for ci, (c0, c1) in enumerate(self.chains):
                st = self.streams[ci]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    self._step_rows(ci, c0, c1, t, t_dev)
            for st in self.streams:
                main.wait_stream(st)
        if t_dev is not None:
            call('pa_decode_advance', t_dev, _stream())

    def _step_rows(self, ci, c0, c1, t, t_dev):
        """Decode step t for the sequences c0..c1-1 (one chain) on the current stream."""
        m = self.m
        B, d = c1 - c0, self.y.shape[1]
        S = self.kpm.shape[1]
        H, dh, T, V = m.num_head, d // m.num_head, m.max_output_length, m.vocab_size
        scale = dh ** -0.5
        e_val = m.input_embeddings['input_value'].weight
        y, other, o = self.y[c0:c1], self.y2[c0:c1], self.o[c0:c1]
        samples, attach = self.samples[c0:c1], self.attach[c0:c1]
        call('pa_decode_embed', samples.data_ptr(), T, B, t, t_dev, m.num_output_dof, e_val.data_ptr(),
             m.query_coord_embedding.weight.data_ptr(), m.query_pos_embedding.weight.data_ptr(), d, y.data_ptr(), _stream())
        for li, l in enumerate(m.decoder.layers):
            sa, ca = l.self_attn, l.multihead_attn
            qkv = self._lin(y, sa.in_proj_weight, sa.in_proj_bias, (ci, 'qkv', li))                 # [B,3d]
            base = qkv.data_ptr()
            call('pa_decode_attn', base, 3 * d, base + 4 * d, base + 8 * d, 3 * d, self.self_k[li][c0:c1].data_ptr(),
                 self.self_v[li][c0:c1].data_ptr(), T, d, t, t + 1, t_dev, None, None, B, H, dh, scale, o.data_ptr(), _stream())
            a = self._lin(o, sa.out_proj.weight, sa.out_proj.bias, (ci, 'so', li))
            y, other = self._add_ln(y, a, l.norm1, m.layer_eps, other), y
            q = self._lin(y, ca.in_proj_weight, ca.in_proj_bias, (ci, 'cq', li), rows=slice(0, d))
            kvb = self.cross_kv[li][c0:c1].data_ptr()
            call('pa_decode_attn', q.data_ptr(), d, None, None, 0, kvb, kvb + 4 * d, S, 2 * d, 0, S, None, self.kpm[c0:c1].data_ptr(),
                 self.kv_len[c0:c1].data_ptr(), B, H, dh, scale, o.data_ptr(), _stream())
            a = self._lin(o, ca.out_proj.weight, ca.out_proj.bias, (ci, 'co', li))
            y, other = self._add_ln(y, a, l.norm2, m.layer_eps, other), y
            h = self._lin(y, l.linear1.weight, l.linear1.bias, (ci, 'f1', li), relu=True)
            f = self._lin(h, l.linear2.weight, l.linear2.bias, (ci, 'f2', li))
            y, other = self._add_ln(y, f, l.norm3, m.layer_eps, other), y
        hfin_t = self._add_ln(y, None, m.decoder.norm, 1e-5, other)
        h3 = self._split(hfin_t, (ci, 'heads')) if self.tc3 else None          # one split feeds both head GEMMs
        lv = self._lin(hfin_t, m.vocab_head.weight, m.vocab_head.bias, (ci, 'lv'), x3=h3)
        pf = self._lin(hfin_t, m.pointer_head.weight, m.pointer_head.bias, (ci, 'pf'), x3=h3)
        sw = self._lin(hfin_t, m.switch_head.weight, m.switch_head.bias, (ci, 'sw'))
        call('pa_decode_head', hfin_t.data_ptr(), lv.data_ptr(), pf.data_ptr(), sw.data_ptr(), self.hfin[c0:c1].data_ptr(), T, B, d,
             V, t, t_dev, m.token.END, samples.data_ptr(), attach.data_ptr(), T, self.first_end[c0:c1].data_ptr(), _stream())

    def _capture(self):
        """Warm up (cuBLAS workspaces, kernel attributes) on a side stream, then capture one step."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):
                self.t_dev.zero_()
                self._step(0, self.t_dev.data_ptr())
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step(0, self.t_dev.data_ptr())
        self.graph = g

    # ------------------------------------------------------------------ whole decode
    @torch.no_grad()
    def run(self, memory, in_kpm):
        m = self.m
        B, S, d = memory.shape
        T = m.max_output_length
        self._buffers(B, S, memory.device)
        # cross-attention K/V: one projection of the encoder memory per layer, kept for all steps
        mem2d = memory.reshape(B * S, d)
        mem3 = ops.split3(mem2d) if ops.GEMM_IMPL == 'tc' else None      # one 3xTF32 split of the memory feeds all layers
        for li, l in enumerate(m.decoder.layers):
            ca = l.multihead_attn
            if mem3 is not None:
                W3 = ops.x3_weight(ca.in_proj_weight)[d:]
                ops.gemm_tf32(mem3, W3, self.cross_kv[li], B * S, 2 * d, 3 * d, lda=3 * d, ldb=3 * d, ldc=2 * d, bias=ca.in_proj_bias[d:])
            else:
                torch.addmm(ca.in_proj_bias[d:], mem2d, ca.in_proj_weight[d:].t(), out=self.cross_kv[li].view(B * S, 2 * d))
        del mem3
        self.kpm.copy_(in_kpm)
        self.kv_len.copy_(ops._kv_len(in_kpm))
        if self.tc3:                                                     # 3xTF32 weight copies follow the current weights
            for l in m.decoder.layers:                                   # (pointer-stable buffers: the captured graph stays valid)
                for W in (l.self_attn.in_proj_weight, l.self_attn.out_proj.weight, l.multihead_attn.in_proj_weight,
                          l.multihead_attn.out_proj.weight, l.linear1.weight, l.linear2.weight):
                    ops.x3_weight(W)
            ops.x3_weight(m.vocab_head.weight)
            ops.x3_weight(m.pointer_head.weight)
        if self.mode == 'fused':
            args, keep = self._fused_args(B, S)
            call('pa_decode_fused', C.byref(args), _stream())
            st = self.state[:4].tolist()                 # the one host sync of the whole decode
            # the reference stops after the step at which every sequence has emitted END (ref models.py:306)
            n_steps = st[3] + 1 if st[1] >= B else T
            del keep
            return self.samples[:, :n_steps].clone(), self.attach[:, :n_steps].clone()
        if self.use_graph and self.graph is None:
            self._capture()
        self.first_end.fill_(T)
        self.samples.zero_()
        self.attach.fill_(-1)
        self.t_dev.zero_()
        n_steps = T
        for t in range(T):
            if self.use_graph:
                self.graph.replay()
            else:
                self._step(t, None)
            if (t + 1) % self.poll == 0 or t == T - 1:
                last = int(self.first_end.max().item())          # the only host sync, every `poll` steps
                if last < T:
                    n_steps = last + 1
                    break
        return self.samples[:, :n_steps].clone(), self.attach[:, :n_steps].clone()
