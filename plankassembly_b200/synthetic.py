"""Synthetic drawings in the exact tensor layout the reference's LineDataset produces.

The reference's data path (ref: plankassembly/datasets/line_data.py:34-109) is out
of scope; its OUTPUT layout is the hot path's INPUT contract, so this module
re-creates that layout from seeded random plank assemblies:

* ``input_value``  int64 [S]: n lines x 4 quantised coords (9-bit, ref:
  datasets/data_utils.py:6-12), lines lexsorted by (view, x1, x2, y1, y2)
  (ref: line_data.py:41-42), then END, then PAD.
* ``input_pos``    line index within its view, repeated x4; 0 on END/PAD.
* ``input_coord``  i % 4; ``input_view`` in {0,1,2}; ``input_type`` in {0,1}.
* ``input_mask``   value == PAD (END is *not* masked).
* ``output_value`` planks x 6 (xmin,ymin,zmin,xmax,ymax,zmax) then END then PAD;
  first plank is the overall bounding box.
* ``output_label`` = value, except attached coordinates = VOCAB + index of the
  earlier output position carrying the same value (must satisfy the pointer
  mask of ref: models.py:91-101).
* ``output_mask``  value == PAD.

Tensor width is MAX_INPUT_LENGTH-1 for the inputs (ref: line_data.py:65-72 pads
the value row one short and the others to match) and MAX_OUTPUT_LENGTH for the
outputs.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np
import torch

END, PAD, VOCAB = 512, 513, 514


class Cfg(dict):
    """Attribute-access dict; stands in for detectron2's CfgNode in tests/bench."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return Cfg(v) if isinstance(v, dict) and not isinstance(v, Cfg) else v


def make_cfg(num_model=512, num_head=8, num_feedforward=1024, dropout=0.2,
             enc_layers=6, dec_layers=6, max_input_length=513, max_output_length=256,
             batch_size=16, lr=1e-4):
    """hparams tree with the keys of ref: configs/train_complete.yaml:25-64."""
    return Cfg({
        'BATCH_SIZE': batch_size, 'LR': lr, 'THRESHOLD': 0.5,
        'DATA': {'NUM_INPUT_DOF': 4, 'NUM_OUTPUT_DOF': 6, 'VOCAB_SIZE': VOCAB, 'NUM_VIEW': 3,
                 'NUM_TYPE': 2, 'MAX_INPUT_LENGTH': max_input_length,
                 'MAX_OUTPUT_LENGTH': max_output_length, 'NUM_BITS': 9},
        'TOKEN': {'END': END, 'PAD': PAD},
        'MODEL': {'NUM_MODEL': num_model, 'NUM_HEAD': num_head, 'NUM_FEEDFORWARD': num_feedforward,
                  'DROPOUT': dropout, 'ACTIVATION': 'relu', 'NORMALIZE_BEFORE': True,
                  'NUM_ENCODER_LAYERS': enc_layers, 'NUM_DECODER_LAYERS': dec_layers},
    })


# The named workloads of BASELINE.json `configs`.
def config1(dropout=0.0):   # d=256, 2+2 layers, B=4, S=1199, T=128 (CPU-runnable)
    return make_cfg(256, 8, 1024, dropout, 2, 2, 1200, 128, batch_size=4)


def config2(dropout=0.2):   # full model, B=64, S=512, T=256
    return make_cfg(512, 8, 1024, dropout, 6, 6, 513, 256, batch_size=64)


def config4(dropout=0.2):   # train_visible.yaml shapes, S=999, T=128
    return make_cfg(512, 8, 1024, dropout, 6, 6, 1000, 128, batch_size=64)


def fixture_cfg(dropout=0.0):   # full model whose tables cover BOTH configs[1] (S=512,T=256) and configs[3] (S=999,T=128) batches
    return make_cfg(512, 8, 1024, dropout, 6, 6, 1000, 256, batch_size=64)


def tiny_cfg(dropout=0.0):  # small committed-fixture model (tests/golden)
    return make_cfg(128, 4, 256, dropout, 2, 2, 300, 64, batch_size=4)


def quantize(v, n_bits=9):
    """[-1,1] floats -> ints in [0, 2^n-1] (same formula as ref: data_utils.py:6-12)."""
    return ((np.asarray(v, dtype=np.float64) + 1.0) * (2 ** n_bits - 1) / 2.0).astype(np.int64)


def pointer_allowed(i, j):
    """6-periodic validity table of ref: models.py:91-101 (rows < 6 point nowhere)."""
    if i < 6 or j >= i:
        return False
    if j < 6:
        return j == i % 6
    return j % 6 == (i % 6 + 3) % 6


def _make_planks(rng, n_planks):
    """Axis-aligned boxes in [-1,1]^3 that share faces, so attachments exist."""
    lo = rng.uniform(-0.95, -0.6, 3)
    hi = rng.uniform(0.6, 0.95, 3)
    planks = [np.concatenate([lo, hi])]
    for _ in range(n_planks - 1):
        box = np.empty(6)
        for ax in range(3):
            cands_lo = [planks[0][ax]] + [p[3 + ax] for p in planks[1:]]
            cands_hi = [planks[0][3 + ax]] + [p[ax] for p in planks[1:]]
            for _try in range(8):
                a = rng.choice(cands_lo) if rng.random() < 0.6 else rng.uniform(lo[ax], hi[ax])
                b = rng.choice(cands_hi) if rng.random() < 0.6 else rng.uniform(lo[ax], hi[ax])
                if b - a > 0.02:
                    break
            else:
                a, b = lo[ax], hi[ax]
            box[ax], box[3 + ax] = a, b
        planks.append(box)
    return np.stack(planks)


def _plank_lines(planks):
    """Three orthographic rectangles per plank -> (x1,y1,x2,y2) edges, view ids."""
    lines, views = [], []
    for p in planks:
        for view, (u, v) in enumerate(((1, 2), (0, 2), (0, 1))):   # drop x / y / z
            u0, u1, v0, v1 = p[u], p[3 + u], p[v], p[3 + v]
            for seg in ((u0, v0, u1, v0), (u0, v1, u1, v1), (u0, v0, u0, v1), (u1, v0, u1, v1)):
                lines.append(seg)
                views.append(view)
    return np.asarray(lines), np.asarray(views)


def _noise(rng, lines, views, types, ratio, length=0.02):
    """Delete-or-shorten a fraction of the lines (behaviour of ref: data_utils.py:24-68)."""
    n = len(lines)
    sel = rng.choice(n, int(math.ceil(n * ratio)), replace=False)
    keep = np.ones(n, dtype=bool)
    lines = lines.copy()
    for i in sel:
        if rng.random() > 0.5:
            keep[i] = False
            continue
        x1, y1, x2, y2 = lines[i]
        ln = math.hypot(x2 - x1, y2 - y1)
        cut = round(rng.random() * length, 3)
        if ln <= cut:
            keep[i] = False
            continue
        f = cut / ln
        if rng.random() > 0.5:
            lines[i] = (x1, y1, x2 - f * (x2 - x1), y2 - f * (y2 - y1))
        else:
            lines[i] = (x1 + f * (x2 - x1), y1 + f * (y2 - y1), x2, y2)
    return lines[keep], views[keep], types[keep]


def make_sample(idx, max_input_length, max_output_length, seed=2022, noise_ratio=0.0, canonical=False):
    """One drawing -> dict of 1-D numpy arrays in LineDataset layout.
    canonical=True lists the planks after the bounding box in lexicographic order of their quantised coordinates (a
    learnable output order, as a curated dataset has; the seeded fixtures of round 1 keep the generation order)."""
    rng = np.random.default_rng(seed + idx)
    S, T = max_input_length - 1, max_output_length
    max_planks = (T - 1) // 6
    n_planks = int(rng.integers(max(2, max_planks // 2), max_planks + 1))
    planks = _make_planks(rng, n_planks)
    if canonical:
        qp = quantize(planks[1:])
        planks = np.concatenate([planks[:1], planks[1:][np.lexsort(qp.T[::-1])]])

    lines, views = _plank_lines(planks)
    q = quantize(lines)
    _, uniq = np.unique(np.concatenate([q, views[:, None]], 1), axis=0, return_index=True)
    uniq.sort()
    lines, views = lines[uniq], views[uniq]
    types = rng.integers(0, 2, len(lines))
    max_lines = (S - 1) // 4
    if len(lines) > max_lines:
        n_keep = int(rng.integers(max_lines // 2, max_lines + 1))
        sel = np.sort(rng.choice(len(lines), n_keep, replace=False))
        lines, views, types = lines[sel], views[sel], types[sel]
    if noise_ratio > 0:
        lines, views, types = _noise(rng, lines, views, types, noise_ratio)

    # ---- input sequence (layout of ref: line_data.py:34-83)
    val = quantize(lines)
    order = np.lexsort(np.concatenate([val, views[:, None]], 1).T[[3, 1, 2, 0, 4]])
    val, views, types = val[order].reshape(-1), views[order], types[order]
    _, counts = np.unique(views, return_counts=True)
    pos = np.concatenate([np.arange(c) for c in counts])
    n_tok = len(val)
    input_value = np.full(S, PAD, dtype=np.int64)
    input_value[:n_tok] = val
    input_value[n_tok] = END

    def padded(a):
        out = np.zeros(S, dtype=np.int64)
        out[:n_tok] = a
        return out

    sample = {
        'input_value': input_value,
        'input_pos': padded(np.repeat(pos, 4)),
        'input_coord': padded(np.arange(n_tok) % 4),
        'input_view': padded(np.repeat(views, 4)),
        'input_type': padded(np.repeat(types, 4)),
        'input_mask': input_value == PAD,
    }

    # ---- output sequence (layout of ref: line_data.py:85-109)
    ov = quantize(planks).reshape(-1)
    n_out = len(ov)
    output_value = np.full(T, PAD, dtype=np.int64)
    output_value[:n_out] = ov
    output_value[n_out] = END
    label = output_value.copy()
    for i in range(6, n_out):           # attach to the first admissible equal-valued position
        for j in range(i):
            if pointer_allowed(i, j) and ov[j] == ov[i]:
                label[i] = VOCAB + j
                break
    sample.update(output_value=output_value, output_label=label, output_mask=output_value == PAD)
    return sample


def make_batch(indices, max_input_length, max_output_length, seed=2022, noise_ratio=0.0,
               device='cpu', with_type=True, canonical=False):
    """Collate samples into the batch dict `PlankModel.forward` consumes."""
    samples = [make_sample(i, max_input_length, max_output_length, seed, noise_ratio, canonical) for i in indices]
    batch = {'name': [f'synthetic_{i:05d}' for i in indices]}
    for k in samples[0]:
        if k == 'input_type' and not with_type:
            continue
        batch[k] = torch.from_numpy(np.stack([s[k] for s in samples])).to(device)
    return batch


def batch_for(cfg, indices, **kw):
    return make_batch(indices, cfg.DATA.MAX_INPUT_LENGTH, cfg.DATA.MAX_OUTPUT_LENGTH, **kw)


def init_state_dict(cfg, seed=2022, dtype=torch.float32):
    """Seeded parameters under the reference's state_dict names.

    Same distributions as ref: models.py:78-83 + torch module defaults (xavier-uniform on
    every tensor with dim>1 incl. embeddings; Linear biases U(+-1/sqrt(fan_in)); attention
    biases 0; LayerNorm 1/0), drawn from our own generator so that fixtures can be re-created
    from a seed anywhere (this container, the GPU box) without the reference present.
    """
    g = torch.Generator().manual_seed(seed)
    d, ff, V = cfg.MODEL.NUM_MODEL, cfg.MODEL.NUM_FEEDFORWARD, cfg.DATA.VOCAB_SIZE
    n_in = math.ceil(cfg.DATA.MAX_INPUT_LENGTH / cfg.DATA.NUM_INPUT_DOF)
    n_out = math.ceil(cfg.DATA.MAX_OUTPUT_LENGTH / cfg.DATA.NUM_OUTPUT_DOF)
    sd = {}

    def xavier(name, *shape):
        bound = math.sqrt(6.0 / (shape[0] + shape[1]))
        sd[name] = (torch.rand(*shape, generator=g, dtype=torch.float64) * 2 - 1).mul_(bound).to(dtype)

    def bias(name, n, fan_in):
        b = 1.0 / math.sqrt(fan_in) if fan_in else 0.0
        sd[name] = (torch.rand(n, generator=g, dtype=torch.float64) * 2 - 1).mul_(b).to(dtype)

    def norm(prefix):
        sd[prefix + '.weight'] = torch.ones(d, dtype=dtype)
        sd[prefix + '.bias'] = torch.zeros(d, dtype=dtype)

    def attn(prefix):
        xavier(prefix + '.in_proj_weight', 3 * d, d)
        sd[prefix + '.in_proj_bias'] = torch.zeros(3 * d, dtype=dtype)
        xavier(prefix + '.out_proj.weight', d, d)
        sd[prefix + '.out_proj.bias'] = torch.zeros(d, dtype=dtype)

    def ffn(prefix):
        xavier(prefix + '.linear1.weight', ff, d)
        bias(prefix + '.linear1.bias', ff, d)
        xavier(prefix + '.linear2.weight', d, ff)
        bias(prefix + '.linear2.bias', d, ff)

    for name, rows in (('input_value', V), ('input_pos', n_in), ('input_coord', cfg.DATA.NUM_INPUT_DOF),
                       ('input_view', cfg.DATA.NUM_VIEW), ('input_type', cfg.DATA.NUM_TYPE)):
        xavier(f'input_embeddings.{name}.weight', rows, d)
    xavier('query_coord_embedding.weight', cfg.DATA.NUM_OUTPUT_DOF, d)
    xavier('query_pos_embedding.weight', n_out, d)
    for i in range(cfg.MODEL.NUM_ENCODER_LAYERS):
        p = f'encoder.layers.{i}'
        attn(p + '.self_attn'); ffn(p); norm(p + '.norm1'); norm(p + '.norm2')
    norm('encoder.norm')
    for i in range(cfg.MODEL.NUM_DECODER_LAYERS):
        p = f'decoder.layers.{i}'
        attn(p + '.self_attn'); attn(p + '.multihead_attn'); ffn(p)
        norm(p + '.norm1'); norm(p + '.norm2'); norm(p + '.norm3')
    norm('decoder.norm')
    xavier('vocab_head.weight', V, d); bias('vocab_head.bias', V, d)
    xavier('pointer_head.weight', d, d); bias('pointer_head.bias', d, d)
    xavier('switch_head.weight', 1, d); bias('switch_head.bias', 1, d)
    return sd


TOKEN = SimpleNamespace(END=END, PAD=PAD)


# ---- compact storage of trained fixture weights: int8 in groups of 32 along the last dim, fp16 scale per group.
# The fixture's weights ARE the dequantised values: the reference (oracle/gen_golden.py) and the CUDA path both load them.
Q_GROUP = 32


def quantize_state_dict(sd):
    """-> dict of numpy arrays: '<name>.q' int8 + '<name>.s' fp16 for 2-D tensors with a last dim % 32 == 0, else '<name>' fp32."""
    out = {}
    for k, v in sd.items():
        v = v.detach().float().cpu()
        if v.dim() == 2 and v.shape[1] % Q_GROUP == 0 and v.numel() >= 4096:
            g = v.view(v.shape[0], -1, Q_GROUP)
            s = (g.abs().amax(-1, keepdim=True) / 127.0).clamp_min(1e-12).half()
            q = torch.round(g / s.float()).clamp_(-127, 127).to(torch.int8)
            out[k + '.q'], out[k + '.s'] = q.view_as(v).numpy(), s.squeeze(-1).numpy()
        else:
            out[k] = v.numpy()
    return out


def dequantize_state_dict(z):
    """Inverse of quantize_state_dict (z: mapping name -> numpy array, e.g. an open .npz)."""
    sd = {}
    for k in z.keys():
        if k.endswith('.q'):
            n = k[:-2]
            q = torch.from_numpy(np.asarray(z[k])).float()
            s = torch.from_numpy(np.asarray(z[n + '.s'])).float()
            sd[n] = (q.view(q.shape[0], -1, Q_GROUP) * s.unsqueeze(-1)).view_as(q).contiguous()
        elif not k.endswith('.s'):
            sd[k] = torch.from_numpy(np.asarray(z[k])).float()
    return sd
