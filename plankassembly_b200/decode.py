"""KV-cached greedy decoder (inference path of PlankModel.eval_step).

The reference regenerates everything on every step (ref models.py:284-307): it re-embeds the
whole prefix, re-runs all decoder layers over all positions, re-projects the encoder memory to
K/V in every layer and builds the full [B,t,V+t] distribution to read its last row, with >= 3
host syncs per step.  Here the per-sequence state is persistent in HBM:

  cross K/V  [L_dec][B,S,2d]   projected ONCE from the encoder memory,
  self  K/V  [L_dec][B,Tmax,d] appended by the attention kernel at each step,
  Hfin       [B,Tmax,d]        final-normed hidden of every emitted position = pointer keys,
  samples / attach [B,Tmax] int64, first_end [B] int32.

Each step touches one new position only; END detection stays on the device and the host looks
at it every `poll` steps.  The reference keeps decoding rows that already emitted END until ALL
rows have (ref models.py:306) and returns those tokens too; the same happens here and the result
is truncated to the reference's stop step max_b(first_end)+1, so `samples`/`attach` are identical.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from ._lib import call


def _stream():
    return torch.cuda.current_stream().cuda_stream


class GreedyDecoder:
    def __init__(self, model, poll=16):
        self.m = model
        self.poll = poll
        self._buf_key = None

    def _buffers(self, B, S, device):
        m = self.m
        key = (B, S, device)
        if self._buf_key == key:
            return
        d, T, L = m.num_model, m.max_output_length, len(m.decoder.layers)
        f32 = dict(device=device, dtype=torch.float32)
        self.self_k = [torch.empty(B, T, d, **f32) for _ in range(L)]
        self.self_v = [torch.empty(B, T, d, **f32) for _ in range(L)]
        self.hfin = torch.empty(B, T, d, **f32)
        self.y = torch.empty(B, d, **f32)
        self.o = torch.empty(B, d, **f32)
        self.samples = torch.empty(B, T, device=device, dtype=torch.int64)
        self.attach = torch.empty(B, T, device=device, dtype=torch.int64)
        self.first_end = torch.empty(B, device=device, dtype=torch.int32)
        self._buf_key = key

    def _add_ln(self, x, a, norm, eps, out):
        call('pa_add_ln_fwd', x.data_ptr(), a.data_ptr() if a is not None else None, norm.weight.data_ptr(),
             norm.bias.data_ptr(), eps, 0.0, 0, 0, x.shape[0], x.shape[1], out.data_ptr(), None, None, None, _stream())
        return out

    @torch.no_grad()
    def run(self, memory, in_kpm):
        m = self.m
        B, S, d = memory.shape
        H, dh, T, V = m.num_head, m.num_model // m.num_head, m.max_output_length, m.vocab_size
        dev = memory.device
        self._buffers(B, S, dev)
        scale = dh ** -0.5
        layers = list(m.decoder.layers)
        # cross-attention K/V: one projection of the encoder memory per layer, kept for all steps
        cross_kv = [F.linear(memory, l.multihead_attn.in_proj_weight[d:], l.multihead_attn.in_proj_bias[d:]) for l in layers]
        self.first_end.fill_(T)
        self.samples.zero_()
        self.attach.fill_(-1)
        e_val = m.input_embeddings['input_value'].weight
        y2 = torch.empty_like(self.y)
        n_steps = T
        for t in range(T):
            y = self.y
            call('pa_decode_embed', self.samples.data_ptr(), T, B, t, m.num_output_dof, e_val.data_ptr(),
                 m.query_coord_embedding.weight.data_ptr(), m.query_pos_embedding.weight.data_ptr(), d, y.data_ptr(), _stream())
            for li, l in enumerate(layers):
                sa, ca = l.self_attn, l.multihead_attn
                qkv = F.linear(y, sa.in_proj_weight, sa.in_proj_bias)                          # [B,3d]
                base = qkv.data_ptr()
                call('pa_decode_attn', base, 3 * d, base + 4 * d, base + 8 * d, 3 * d, self.self_k[li].data_ptr(),
                     self.self_v[li].data_ptr(), T, d, t, t + 1, None, B, H, dh, scale, self.o.data_ptr(), _stream())
                a = F.linear(self.o, sa.out_proj.weight, sa.out_proj.bias)
                y = self._add_ln(y, a, l.norm1, m.layer_eps, y2 if y is self.y else self.y)
                q = F.linear(y, ca.in_proj_weight[:d], ca.in_proj_bias[:d])
                kvb = cross_kv[li].data_ptr()
                call('pa_decode_attn', q.data_ptr(), d, None, None, 0, kvb, kvb + 4 * d, S, 2 * d, 0, S, in_kpm.data_ptr(),
                     B, H, dh, scale, self.o.data_ptr(), _stream())
                a = F.linear(self.o, ca.out_proj.weight, ca.out_proj.bias)
                y = self._add_ln(y, a, l.norm2, m.layer_eps, y2 if y is self.y else self.y)
                h = torch.relu_(F.linear(y, l.linear1.weight, l.linear1.bias))
                f = F.linear(h, l.linear2.weight, l.linear2.bias)
                y = self._add_ln(y, f, l.norm3, m.layer_eps, y2 if y is self.y else self.y)
            hfin_t = self._add_ln(y, None, m.decoder.norm, 1e-5, y2 if y is self.y else self.y)
            lv = F.linear(hfin_t, m.vocab_head.weight, m.vocab_head.bias)
            pf = F.linear(hfin_t, m.pointer_head.weight, m.pointer_head.bias)
            sw = F.linear(hfin_t, m.switch_head.weight, m.switch_head.bias)
            call('pa_decode_head', hfin_t.data_ptr(), lv.data_ptr(), pf.data_ptr(), sw.data_ptr(), self.hfin.data_ptr(), T, B, d,
                 V, t, m.token.END, self.samples.data_ptr(), self.attach.data_ptr(), T, self.first_end.data_ptr(), _stream())
            if (t + 1) % self.poll == 0 or t == T - 1:
                last = int(self.first_end.max().item())          # the only host sync, every `poll` steps
                if last < T:
                    n_steps = last + 1
                    break
        return self.samples[:, :n_steps].clone(), self.attach[:, :n_steps].clone()
