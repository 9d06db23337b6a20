"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by plankassembly_b200/.

CPU restatement (explicit fp32/fp64 torch math, no nn.Transformer, no fused
attention) of the reference hot path ``/root/reference/plankassembly/models.py``.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm
may import this file, and only as the checker or as the timed CPU baseline.

Parity pin: the reference ships no tests or golden vectors.
This restatement is pinned against outputs of the reference itself, produced in the
dev container by ``oracle/gen_golden.py`` (imports /root/reference) and committed
under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.

The arithmetic of the reference lives in PyTorch (pinned 1.10.0 in the reference's
environment.yml:95, 2.11.0 here): nn.TransformerEncoder/Decoder in post-norm mode.
Semantics restated from the reference's call sites:

* models.py:60-69   layer construction.  ``normalize_before`` lands in the
  ``layer_norm_eps`` slot => per-layer LayerNorm eps == 1.0 and post-norm layers;
  the two final norms keep eps = 1e-5.
* models.py:103-138 embeddings;  :140-188 distributions;  :190-233 train_step;
  :235-256 sampling;  :267-323 greedy loop.
* torch nn/functional.py multi_head_attention_forward: packed in-proj, heads are
  contiguous dh-slices, softmax(q k^T / sqrt(dh) + additive -inf masks), out-proj.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

NEG_INF = float('-inf')


class OraclePlankModel:
    """Functional model over a dict of tensors keyed by the reference's state_dict names."""

    def __init__(self, cfg, state_dict, dtype=torch.float32, requires_grad=False):
        self.d = cfg.MODEL.NUM_MODEL
        self.H = cfg.MODEL.NUM_HEAD
        self.p_drop = float(cfg.MODEL.DROPOUT)
        self.n_enc = cfg.MODEL.NUM_ENCODER_LAYERS
        self.n_dec = cfg.MODEL.NUM_DECODER_LAYERS
        self.V = cfg.DATA.VOCAB_SIZE
        self.dof = cfg.DATA.NUM_OUTPUT_DOF
        self.max_out = cfg.DATA.MAX_OUTPUT_LENGTH
        self.END, self.PAD = cfg.TOKEN.END, cfg.TOKEN.PAD
        self.P = {k: v.detach().clone().to(dtype).requires_grad_(requires_grad) for k, v in state_dict.items()}
        self.training = False

    def parameters(self):
        return list(self.P.values())

    # ------------------------------------------------------------------ building blocks
    def _drop(self, x):
        return F.dropout(x, self.p_drop, True) if (self.training and self.p_drop > 0) else x

    def _ln(self, x, prefix, eps):
        mu = x.mean(-1, keepdim=True)
        var = ((x - mu) ** 2).mean(-1, keepdim=True)          # biased variance
        return (x - mu) / torch.sqrt(var + eps) * self.P[prefix + '.weight'] + self.P[prefix + '.bias']

    def _mha(self, prefix, xq, xkv, add_mask):
        """add_mask broadcastable to [B,H,Lq,Lk], entries 0 / -inf."""
        d, H = self.d, self.H
        dh = d // H
        W, b = self.P[prefix + '.in_proj_weight'], self.P[prefix + '.in_proj_bias']
        q = xq @ W[:d].T + b[:d]
        k = xkv @ W[d:2 * d].T + b[d:2 * d]
        v = xkv @ W[2 * d:].T + b[2 * d:]
        B, Lq, Lk = xq.shape[0], xq.shape[1], xkv.shape[1]
        q = q.view(B, Lq, H, dh).transpose(1, 2)
        k = k.view(B, Lk, H, dh).transpose(1, 2)
        v = v.view(B, Lk, H, dh).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
        if add_mask is not None:
            s = s + add_mask
        p = self._drop(torch.softmax(s, dim=-1))
        o = (p @ v).transpose(1, 2).reshape(B, Lq, d)
        return o @ self.P[prefix + '.out_proj.weight'].T + self.P[prefix + '.out_proj.bias']

    def _ffn(self, prefix, x):
        h = self._drop(torch.relu(x @ self.P[prefix + '.linear1.weight'].T + self.P[prefix + '.linear1.bias']))
        return h @ self.P[prefix + '.linear2.weight'].T + self.P[prefix + '.linear2.bias']

    @staticmethod
    def _kpm(mask, dtype):
        """bool [B,L] (True = PAD key) -> additive [B,1,1,L]."""
        return torch.zeros(mask.shape, dtype=dtype).masked_fill_(mask, NEG_INF)[:, None, None, :]

    # ------------------------------------------------------------------ models.py:103-138
    def embed_input(self, batch):
        x = 0
        for key, ids in batch.items():
            if key[:5] != 'input' or 'mask' in key:
                continue
            x = x + self.P[f'input_embeddings.{key}.weight'][ids]
        return x

    def embed_output(self, output):
        B, n = output.shape
        t = torch.arange(n)
        e = (self.P['input_embeddings.input_value.weight'][output]
             + self.P['query_coord_embedding.weight'][t % self.dof][None]
             + self.P['query_pos_embedding.weight'][t // self.dof][None])
        return torch.cat([torch.zeros(B, 1, self.d, dtype=e.dtype), e], dim=1)

    # ------------------------------------------------------------------ encoder / decoder
    def encode(self, x, input_mask):
        m = self._kpm(input_mask, x.dtype)
        for i in range(self.n_enc):
            p = f'encoder.layers.{i}'
            x = self._ln(x + self._drop(self._mha(p + '.self_attn', x, x, m)), p + '.norm1', 1.0)
            x = self._ln(x + self._drop(self._ffn(p, x)), p + '.norm2', 1.0)
        return self._ln(x, 'encoder.norm', 1e-5)

    def decode(self, y, memory, input_mask, tgt_kpm=None):
        T = y.shape[1]
        self_mask = torch.triu(torch.full((T, T), NEG_INF, dtype=y.dtype), 1)[None, None]
        if tgt_kpm is not None:
            self_mask = self_mask + self._kpm(tgt_kpm, y.dtype)
        mem_mask = self._kpm(input_mask, y.dtype)
        for i in range(self.n_dec):
            p = f'decoder.layers.{i}'
            y = self._ln(y + self._drop(self._mha(p + '.self_attn', y, y, self_mask)), p + '.norm1', 1.0)
            y = self._ln(y + self._drop(self._mha(p + '.multihead_attn', y, memory, mem_mask)), p + '.norm2', 1.0)
            y = self._ln(y + self._drop(self._ffn(p, y)), p + '.norm3', 1.0)
        return self._ln(y, 'decoder.norm', 1e-5)

    # ------------------------------------------------------------------ models.py:140-188
    def heads(self, h):
        lv = h @ self.P['vocab_head.weight'].T + self.P['vocab_head.bias']
        pf = h @ self.P['pointer_head.weight'].T + self.P['pointer_head.bias']
        lp = (pf @ h.transpose(1, 2)) / self.d
        pi = torch.sigmoid(h @ self.P['switch_head.weight'].T + self.P['switch_head.bias'])
        return lv, lp, pi

    def pointer_mask(self, sz):
        i = torch.arange(sz)[:, None]
        j = torch.arange(sz)[None, :]
        ok = torch.where(j < 6, j == i % 6, j % 6 == (i % 6 + 3) % 6) & (i >= 6)
        return ok

    def dist_train(self, h, eps=1e-6):
        lv, lp, pi = self.heads(h)
        T = h.shape[1]
        upper = torch.triu(torch.ones(T, T, dtype=torch.bool))            # j >= i
        lp = lp.masked_fill(upper[None], eps)                             # a VALUE, not -inf
        dv = torch.log_softmax(lv, -1) + torch.log(torch.clamp(1 - pi, min=eps))
        dp = torch.log_softmax(lp, -1) + torch.log(torch.clamp(pi, min=eps))
        return torch.cat([dv, dp], -1)

    def dist_eval(self, h, eps=1e-6):
        lv, lp, pi = self.heads(h)
        sz = h.shape[1]
        pv = torch.softmax(lv, -1)
        if sz < 6:
            return pv
        upper = torch.triu(torch.ones(sz, sz, dtype=torch.bool))
        pp = torch.softmax(lp.masked_fill(upper[None], NEG_INF), -1) * pi   # row 0 -> NaN, overwritten below
        pv = pv * (1 - pi)
        pp = pp.masked_fill(~self.pointer_mask(sz)[None], eps)
        return torch.cat([pv, pp], -1)

    # ------------------------------------------------------------------ models.py:190-233
    def train_step(self, batch, return_dists=False):
        x = self.embed_input(batch)
        y = self.embed_output(batch['output_value'][:, :-1])
        memory = self.encode(x, batch['input_mask'])
        h = self.decode(y, memory, batch['input_mask'], tgt_kpm=batch['output_mask'])
        dists = self.dist_train(h)
        label = batch['output_label']
        valid = label != self.PAD
        picked = dists.gather(-1, label.clamp(max=dists.shape[-1] - 1)[..., None])[..., 0]
        loss = -(picked * valid).sum() / valid.sum()
        predict = dists.argmax(-1)
        accuracy = (valid & (predict == label)).sum().to(torch.float32) / (valid.sum() + 1e-10)
        out = {'loss': loss, 'accuracy': accuracy}
        if return_dists:
            out.update(dists=dists, hiddens=h, memory=memory)
        return out

    # ------------------------------------------------------------------ models.py:235-323
    def _sample(self, last_row, output):
        c = last_row.argmax(-1)
        attach = torch.where(c >= self.V, c - self.V, torch.full_like(c, -1))
        if output.shape[1] > 0:
            copied = output.gather(1, attach.clamp(min=0)[:, None])[:, 0]
            token = torch.where(c >= self.V, copied, c)
        else:
            token = c
        return token, attach

    def parse_sequence(self, seq):
        keep = torch.cumsum(seq == self.END, 0) == 0
        v = seq[keep]
        n = len(v) // self.dof
        return v[:n * self.dof].reshape(-1, self.dof)

    def _finish(self, batch, output, attach):
        B = output.shape[0]
        return {'samples': output, 'attach': attach,
                'predicts': [self.parse_sequence(output[i]) for i in range(B)],
                'groundtruths': [self.parse_sequence(batch['output_value'][i]) for i in range(B)]}

    @torch.no_grad()
    def eval_step_full(self, batch, return_margins=False):
        """Faithful O(T^3) loop of models.py:284-307: everything recomputed each step."""
        memory = self.encode(self.embed_input(batch), batch['input_mask'])
        B = memory.shape[0]
        output = torch.empty(B, 0, dtype=torch.long)
        attach = torch.empty(B, 0, dtype=torch.long)
        margins = []
        for _ in range(self.max_out):
            h = self.decode(self.embed_output(output), memory, batch['input_mask'])
            last = self.dist_eval(h)[:, -1]
            if return_margins:
                top2 = last.topk(2, -1).values
                margins.append((top2[:, 0] - top2[:, 1]) / top2[:, 0])
            tok, att = self._sample(last, output)
            output = torch.cat([output, tok[:, None]], 1)
            attach = torch.cat([attach, att[:, None]], 1)
            if bool(torch.all(torch.any(output == self.END, dim=1))):
                break
        out = self._finish(batch, output, attach)
        if return_margins:
            out['margins'] = torch.stack(margins, 1)
        return out

    @torch.no_grad()
    def eval_step_cached(self, batch):
        """Incremental decode with self/cross K/V caches and cached final hiddens as pointer
        keys.  Mathematically equal to eval_step_full; this is the
        algorithm the CUDA decode path implements, restated on CPU for small cases."""
        d, H = self.d, self.H
        dh = d // H
        P = self.P
        memory = self.encode(self.embed_input(batch), batch['input_mask'])
        B, S, _ = memory.shape
        mem_mask = self._kpm(batch['input_mask'], memory.dtype)            # [B,1,1,S]
        ck, cv = [], []
        for i in range(self.n_dec):
            W, b = P[f'decoder.layers.{i}.multihead_attn.in_proj_weight'], P[f'decoder.layers.{i}.multihead_attn.in_proj_bias']
            ck.append((memory @ W[d:2 * d].T + b[d:2 * d]).view(B, S, H, dh).transpose(1, 2))
            cv.append((memory @ W[2 * d:].T + b[2 * d:]).view(B, S, H, dh).transpose(1, 2))
        sk = [torch.zeros(B, H, 0, dh, dtype=memory.dtype) for _ in range(self.n_dec)]
        sv = [torch.zeros(B, H, 0, dh, dtype=memory.dtype) for _ in range(self.n_dec)]
        hfin = torch.zeros(B, 0, d, dtype=memory.dtype)
        output = torch.empty(B, 0, dtype=torch.long)
        attach = torch.empty(B, 0, dtype=torch.long)

        def attend(q, k, v, mask):
            s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
            if mask is not None:
                s = s + mask
            return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B, 1, d)

        for t in range(self.max_out):
            if t == 0:
                y = torch.zeros(B, 1, d, dtype=memory.dtype)
            else:
                y = (P['input_embeddings.input_value.weight'][output[:, t - 1]]
                     + P['query_coord_embedding.weight'][(t - 1) % self.dof]
                     + P['query_pos_embedding.weight'][(t - 1) // self.dof])[:, None]
            for i in range(self.n_dec):
                p = f'decoder.layers.{i}'
                W, b = P[p + '.self_attn.in_proj_weight'], P[p + '.self_attn.in_proj_bias']
                qkv = (y @ W.T + b).view(B, 1, 3, H, dh)
                q, k, v = (qkv[:, :, j].transpose(1, 2) for j in range(3))
                sk[i] = torch.cat([sk[i], k], 2)
                sv[i] = torch.cat([sv[i], v], 2)
                a = attend(q, sk[i], sv[i], None) @ P[p + '.self_attn.out_proj.weight'].T + P[p + '.self_attn.out_proj.bias']
                y = self._ln(y + a, p + '.norm1', 1.0)
                W, b = P[p + '.multihead_attn.in_proj_weight'], P[p + '.multihead_attn.in_proj_bias']
                q = (y @ W[:d].T + b[:d]).view(B, 1, H, dh).transpose(1, 2)
                a = attend(q, ck[i], cv[i], mem_mask) @ P[p + '.multihead_attn.out_proj.weight'].T + P[p + '.multihead_attn.out_proj.bias']
                y = self._ln(y + a, p + '.norm2', 1.0)
                y = self._ln(y + self._ffn(p, y), p + '.norm3', 1.0)
            h = self._ln(y, 'decoder.norm', 1e-5)                           # [B,1,d]
            hfin = torch.cat([hfin, h], 1)
            sz = t + 1
            lv = h[:, 0] @ P['vocab_head.weight'].T + P['vocab_head.bias']
            pv = torch.softmax(lv, -1)
            if sz < 6:
                last = pv
            else:
                pi = torch.sigmoid(h[:, 0] @ P['switch_head.weight'].T + P['switch_head.bias'])   # [B,1]
                pf = h[:, 0] @ P['pointer_head.weight'].T + P['pointer_head.bias']
                lp = torch.einsum('bd,bjd->bj', pf, hfin) / d                # [B,sz]; column sz-1 is self
                lp[:, sz - 1] = NEG_INF
                pp = torch.softmax(lp, -1) * pi
                pp = pp.masked_fill(~self.pointer_mask(sz)[sz - 1][None], 1e-6)
                last = torch.cat([pv * (1 - pi), pp], -1)
            tok, att = self._sample(last, output)
            output = torch.cat([output, tok[:, None]], 1)
            attach = torch.cat([attach, att[:, None]], 1)
            if bool(torch.all(torch.any(output == self.END, dim=1))):
                break
        return self._finish(batch, output, attach)


def adam_train_step(model: OraclePlankModel, opt, batch):
    """One optimiser step as Trainer.training_step + configure_optimizers would run it
    (ref: trainer_complete.py:63-71,127-129): forward, backward, Adam."""
    model.training = True
    opt.zero_grad(set_to_none=True)
    out = model.train_step(batch)
    out['loss'].backward()
    opt.step()
    return out
