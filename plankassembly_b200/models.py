"""B200-native drop-in for the reference's ``plankassembly/models.py``.

Same public surface (ref models.py:11-343): ``build_model(cfg)``, ``PlankModel.forward(batch)``
returning ``{'loss','accuracy'}`` in train mode and ``{'samples','attach','predicts',
'groundtruths'}`` in eval mode, ``train_step`` / ``eval_step`` / ``parse_sequence``, the attributes
the Lightning module touches (``token``, ``vocab_size``, ``max_output_length``, ``num_output_dof``)
and the reference's exact ``state_dict`` names/shapes, so its checkpoints load unchanged.

What differs is underneath: no nn.Transformer.  Every op on the path is a hand-written sm_100a
kernel behind the C ABI of include/plank_b200.h (embedding gather-sum, attention, residual+dropout
+LayerNorm, pointer-head distribution/loss, and the dense projections as a tcgen05 TF32 GEMM), and eval mode runs a KV-cached greedy decoder (``decode.py``) instead of the
reference's O(T^3) recompute loop.  There is no CPU path: tensors must live on a B200.

Reproduced quirks: the reference passes ``normalize_before`` into the
``layer_norm_eps`` slot of nn.Transformer*Layer (ref models.py:60-61,66-67), so layers are
POST-norm with eps = float(normalize_before) (1.0 for the shipped configs) while the two final
norms use 1e-5; the pointer head scores against the decoder's own hidden states and divides by d.
"""
from __future__ import annotations

import math
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops, postprocess
from .decode import GreedyDecoder


class _AttnParams(nn.Module):
    """Parameter container named like nn.MultiheadAttention (in_proj_weight/bias, out_proj.*)."""

    def __init__(self, d):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = nn.Linear(d, d)
        nn.init.zeros_(self.out_proj.bias)


class _EncoderLayerParams(nn.Module):
    def __init__(self, d, ff):
        super().__init__()
        self.self_attn = _AttnParams(d)
        self.linear1 = nn.Linear(d, ff)
        self.linear2 = nn.Linear(ff, d)
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = nn.LayerNorm(d)


class _DecoderLayerParams(nn.Module):
    def __init__(self, d, ff):
        super().__init__()
        self.self_attn = _AttnParams(d)
        self.multihead_attn = _AttnParams(d)
        self.linear1 = nn.Linear(d, ff)
        self.linear2 = nn.Linear(ff, d)
        self.norm1 = nn.LayerNorm(d)
        self.norm2 = nn.LayerNorm(d)
        self.norm3 = nn.LayerNorm(d)


class _Stack(nn.Module):
    def __init__(self, layers, norm):
        super().__init__()
        self.layers = nn.ModuleList(layers)
        self.norm = norm


class PlankModel(nn.Module):

    def __init__(self, num_model=512, num_head=8, num_feedforward=1024, dropout=0.1, activation="relu",
                 normalize_before=True, num_encoder_layers=6, num_decoder_layers=6, num_view=3, num_type=2,
                 num_input_dof=4, num_output_dof=6, max_input_length=400, max_output_length=128,
                 vocab_size=514, token=None):
        super().__init__()
        if activation != 'relu':
            raise NotImplementedError('only the relu FFN of the shipped configs is implemented')
        if num_model % 128 or (num_model // num_head) not in (32, 64):
            raise NotImplementedError('kernels are built for d % 128 == 0 and head dim 32 or 64')
        max_num_input = math.ceil(max_input_length / num_input_dof)
        max_num_output = math.ceil(max_output_length / num_output_dof)
        self.num_model, self.num_head = num_model, num_head
        self.max_num_input, self.max_input_length = max_num_input, max_input_length
        self.max_output_length, self.max_num_output = max_output_length, max_num_output
        self.num_input_dof, self.num_output_dof = num_input_dof, num_output_dof
        self.vocab_size = vocab_size
        self.token = token
        self.dropout = float(dropout)
        # the positional-argument quirk of ref models.py:60-61: eps := normalize_before, post-norm
        self.layer_eps = float(normalize_before)
        # 'simt' = fp32 CUDA-core attention (exact), 'tc' = tcgen05 TF32 attention forward
        self.attn_impl = os.environ.get('PLANK_B200_ATTN', 'tc')

        self.input_embeddings = nn.ModuleDict({
            'input_value': nn.Embedding(vocab_size, num_model),
            'input_pos': nn.Embedding(max_num_input, num_model),
            'input_coord': nn.Embedding(num_input_dof, num_model),
            'input_view': nn.Embedding(num_view, num_model),
            'input_type': nn.Embedding(num_type, num_model),
        })
        self.query_coord_embedding = nn.Embedding(num_output_dof, num_model)
        self.query_pos_embedding = nn.Embedding(max_num_output, num_model)
        self.encoder = _Stack([_EncoderLayerParams(num_model, num_feedforward) for _ in range(num_encoder_layers)],
                              nn.LayerNorm(num_model) if normalize_before else None)
        self.decoder = _Stack([_DecoderLayerParams(num_model, num_feedforward) for _ in range(num_decoder_layers)],
                              nn.LayerNorm(num_model))
        self.vocab_head = nn.Linear(num_model, vocab_size)
        self.pointer_head = nn.Linear(num_model, num_model)
        self.switch_head = nn.Linear(num_model, 1)
        self._reset_parameters()
        self._decoder_engine = None

    def _reset_parameters(self):
        """Xavier-uniform on every tensor with dim > 1, embeddings included (ref models.py:78-83)."""
        for p in self.parameters():
            if p.dim() > 1:
                nn.init.xavier_uniform_(p)

    # ------------------------------------------------------------------ building blocks
    # Precision policy.  Training runs the dense contractions (projections, FFN, heads) on the tcgen05
    # tensor cores in TF32 with FP32 accumulation; every tensor-core operand is rounded to nearest TF32
    # by the kernel that produces it (tensors named *_r below), which keeps loss/logits inside the 1e-3
    # parity bar.  Inference (eval_step) keeps everything in exact fp32 so greedy tokens stay bit-exact.
    def _p(self):
        return self.dropout if self.training else 0.0

    def _tf32(self):
        return self.training and ops.GEMM_IMPL == 'tc'

    def _impl(self):
        # tensor-core attention belongs to the TF32 training path; inference stays on the exact kernels
        return ops.ATTN_IMPL[self.attn_impl] if self._tf32() else 0

    @staticmethod
    def _kpm(mask):
        return mask.contiguous().view(torch.uint8)

    def _embed_input(self, inputs, want_r=False):
        keys = [k for k in inputs if 'mask' not in k]
        ids = [inputs[k] for k in keys]
        tables = [self.input_embeddings[k].weight for k in keys]
        return ops.EmbedInput.apply(len(keys), want_r, *ids, *tables)

    def _embed_output(self, output_value, T, want_r=False):
        """Embeds output_value[:, :T-1] behind a zero row -> [B,T,d] (ref models.py:114-138)."""
        return ops.EmbedOutput.apply(output_value, T, self.num_output_dof, want_r, self.input_embeddings['input_value'].weight,
                                     self.query_coord_embedding.weight, self.query_pos_embedding.weight)

    def _add_ln(self, x, a, norm, eps, p, tf, a_bias=None):
        """-> (y, y_r): y_r is the TF32-rounded copy for the next GEMM (y itself in exact mode).
        a_bias: bias of the linear that produced `a` when that GEMM ran bias-free (TF32 path)."""
        if tf:
            return ops.AddLayerNorm.apply(x, a, a_bias, norm.weight, norm.bias, eps, p, True, a is not None)
        y = ops.AddLayerNorm.apply(x, a, None, norm.weight, norm.bias, eps, p, False, False)
        return y, y

    def _ffn(self, layer, x_r, tf):
        ff, d = layer.linear1.weight.shape
        if tf and ops.FFN_FUSED and ff % 32 == 0:
            # both projections and the activation as one autograd node (fused activation backward, ops.FFN)
            return ops.FFN.apply(x_r, layer.linear1.weight, layer.linear1.bias, ops.tf32_weight(layer.linear1.weight),
                                 layer.linear2.weight, ops.tf32_weight(layer.linear2.weight), self._p())
        h = ops.linear(x_r, layer.linear1.weight, layer.linear1.bias, relu=True, p_drop=self._p(), tf32=tf, round_out=True)
        # TF32 path: linear2 runs bias-free, its bias is folded into the following residual+LayerNorm kernel
        return ops.linear(h, layer.linear2.weight, None if tf else layer.linear2.bias, tf32=tf, round_dx=True)

    def _encode(self, x, x_r, in_kpm):
        p, H, tf = self._p(), self.num_head, self._tf32()
        for layer in self.encoder.layers:
            sa = layer.self_attn
            qkv = ops.linear(x_r, sa.in_proj_weight, sa.in_proj_bias, tf32=tf, round_out=True, bias_grad=False)
            a = ops.SelfAttention.apply(qkv, sa.in_proj_bias if tf else None, in_kpm, H, False, p, self._impl(), tf)
            a = ops.linear(a, sa.out_proj.weight, None if tf else sa.out_proj.bias, tf32=tf, round_dx=True)
            x, x_r = self._add_ln(x, a, layer.norm1, self.layer_eps, p, tf, sa.out_proj.bias)
            x, x_r = self._add_ln(x, self._ffn(layer, x_r, tf), layer.norm2, self.layer_eps, p, tf, layer.linear2.bias)
        if self.encoder.norm is not None:
            x, x_r = self._add_ln(x, None, self.encoder.norm, 1e-5, 0.0, tf)
        return x, x_r

    def _decode_train(self, y, y_r, memory_r, in_kpm, out_kpm):
        p, H, d, tf = self._p(), self.num_head, self.num_model, self._tf32()
        mem_acc = ops.DxAccum() if (tf and ops.DX_ACCUM) else None      # the memory's gradient: summed inside the dX GEMM epilogues
        for layer in self.decoder.layers:
            sa, ca = layer.self_attn, layer.multihead_attn
            qkv = ops.linear(y_r, sa.in_proj_weight, sa.in_proj_bias, tf32=tf, round_out=True, bias_grad=False)
            a = ops.SelfAttention.apply(qkv, sa.in_proj_bias if tf else None, out_kpm, H, True, p, self._impl(), tf)
            a = ops.linear(a, sa.out_proj.weight, None if tf else sa.out_proj.bias, tf32=tf, round_dx=True)
            y, y_r = self._add_ln(y, a, layer.norm1, self.layer_eps, p, tf, sa.out_proj.bias)
            w_acc = ops.DxAccum() if tf else None           # one in-projection weight gradient for the q rows and the k/v rows
            q = ops.linear(y_r, ca.in_proj_weight, ca.in_proj_bias, rows=slice(0, d), tf32=tf, round_out=True, bias_grad=False, w_accum=w_acc)
            kv = ops.linear(memory_r, ca.in_proj_weight, ca.in_proj_bias, rows=slice(d, 3 * d), tf32=tf, round_out=True, bias_grad=False,
                            dx_accum=mem_acc, w_accum=w_acc)
            a = ops.CrossAttention.apply(q, kv, ca.in_proj_bias if tf else None, in_kpm, H, p, self._impl(), tf)
            a = ops.linear(a, ca.out_proj.weight, None if tf else ca.out_proj.bias, tf32=tf, round_dx=True)
            y, y_r = self._add_ln(y, a, layer.norm2, self.layer_eps, p, tf, ca.out_proj.bias)
            y, y_r = self._add_ln(y, self._ffn(layer, y_r, tf), layer.norm3, self.layer_eps, p, tf, layer.linear2.bias)
        return self._add_ln(y, None, self.decoder.norm, 1e-5, 0.0, tf)

    def _heads(self, h, h_r):
        tf = self._tf32()
        lv = ops.linear(h_r, self.vocab_head.weight, self.vocab_head.bias, tf32=tf)
        pf = ops.linear(h_r, self.pointer_head.weight, self.pointer_head.bias, tf32=tf, round_out=True)
        lp = ops.pointer_scores(pf, h_r, tf)                       # raw scores; 1/d applied in the kernel
        sw = F.linear(h, self.switch_head.weight, self.switch_head.bias).squeeze(-1)
        return lv, lp, sw

    # ------------------------------------------------------------------ ref models.py:190-233
    def train_step(self, batch, return_dists=False):
        inputs = {k: v for k, v in batch.items() if k[:5] == 'input'}
        in_kpm = self._kpm(batch['input_mask'])
        out_kpm = self._kpm(batch['output_mask'])
        output_value, output_label = batch['output_value'], batch['output_label']
        T = output_value.shape[1]
        tf = self._tf32()
        ops.begin_step()                                       # derived weight copies are rebuilt once per step
        if torch.is_grad_enabled():
            ops.zero_pool.begin_step(output_value.device)      # one memset for all zero-initialised gradient buffers

        if tf and self._p() > 0 and self._impl() == 1:
            # all attention-dropout bit planes of this step, generated on a side stream in the shadow of the main stream
            B, S, H = output_value.shape[0], batch['input_mask'].shape[1], self.num_head
            ops.MASKS.plan([(B, H, S, S)] * len(self.encoder.layers) + [(B, H, T, T), (B, H, T, S)] * len(self.decoder.layers),
                           self._p(), output_value.device)
        if tf:
            x, x_r = self._embed_input(inputs, True)
            y, y_r = self._embed_output(output_value, T, True)
        else:
            x = x_r = self._embed_input(inputs)
            y = y_r = self._embed_output(output_value, T)
        memory, memory_r = self._encode(x, x_r, in_kpm)
        hiddens, hiddens_r = self._decode_train(y, y_r, memory_r, in_kpm, out_kpm)
        lv, lp, sw = self._heads(hiddens, hiddens_r)
        loss, accuracy, predict = ops.DistLoss.apply(lv, lp, sw, output_label, self.token.PAD, 1.0 / self.num_model, tf)
        rets = {'loss': loss, 'accuracy': accuracy}
        if return_dists:                                           # parity tests only
            rets.update(dists=ops.dist_train_full(lv, lp, sw, 1.0 / self.num_model), hiddens=hiddens, memory=memory,
                        predict=predict)
        return rets

    # ------------------------------------------------------------------ ref models.py:258-323
    def parse_sequence(self, sequence):
        valid_mask = torch.cumsum(sequence == self.token.END, 0) == 0
        valid_seq = sequence[valid_mask]
        num_plank = len(valid_seq) // self.num_output_dof
        return valid_seq[:num_plank * self.num_output_dof].reshape(-1, self.num_output_dof)

    @torch.no_grad()
    def eval_step(self, batch):
        """Greedy decoding with persistent K/V caches (same tokens as ref models.py:267-323)."""
        inputs = {k: v for k, v in batch.items() if k[:5] == 'input'}
        in_kpm = self._kpm(batch['input_mask'])
        ops.begin_step()
        x = self._embed_input(inputs)
        memory, _ = self._encode(x, x, in_kpm)
        if self._decoder_engine is None:
            self._decoder_engine = GreedyDecoder(self)
        output, attach = self._decoder_engine.run(memory, in_kpm)
        # ref models.py:309-315 parses every sequence on its own (twice per sample); here one launch per token tensor
        # (postprocess.py) and the lists are views into the two padded plank tensors
        dof, END = self.num_output_dof, self.token.END
        predicts = postprocess.plank_lists(*postprocess.parse_batch(output, END, dof)[:2])
        groundtruths = postprocess.plank_lists(*postprocess.parse_batch(batch['output_value'], END, dof)[:2])
        return {'samples': output, 'attach': attach, 'predicts': predicts, 'groundtruths': groundtruths}

    def forward(self, batch):
        return self.train_step(batch) if self.training else self.eval_step(batch)


def build_model(cfg):
    """Same positional field order as ref models.py:333-343."""
    return PlankModel(
        cfg.MODEL.NUM_MODEL, cfg.MODEL.NUM_HEAD,
        cfg.MODEL.NUM_FEEDFORWARD, cfg.MODEL.DROPOUT,
        cfg.MODEL.ACTIVATION, cfg.MODEL.NORMALIZE_BEFORE,
        cfg.MODEL.NUM_ENCODER_LAYERS, cfg.MODEL.NUM_DECODER_LAYERS,
        cfg.DATA.NUM_VIEW, cfg.DATA.NUM_TYPE,
        cfg.DATA.NUM_INPUT_DOF, cfg.DATA.NUM_OUTPUT_DOF,
        cfg.DATA.MAX_INPUT_LENGTH, cfg.DATA.MAX_OUTPUT_LENGTH,
        cfg.DATA.VOCAB_SIZE, cfg.TOKEN)
