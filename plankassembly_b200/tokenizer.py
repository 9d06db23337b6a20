"""batched, device-side tokeniser with varlen input.

`tokenize_lines` / `tokenize_planks` produce, for a whole batch at once, exactly the tensors
`LineDataset.prepare_input_sequence` / `prepare_output_sequence` build per sample on the CPU (ref:
plankassembly/datasets/line_data.py:34-83, 85-109) -- same quantisation (fp64, truncation), same stable lexicographic line
order, same END / PAD layout -- plus `kv_len`, the valid length per drawing.  A loader that ships raw geometry in varlen form
(all lines of the batch concatenated + offsets) moves 48 B per line to the device instead of 6 padded int64 planes
(1199 x 8 B x 5 per drawing at MAX_INPUT_LENGTH 1200).  The reference's dataset classes stay untouched (north_star); this is
the B200-side entry point for loaders that want it.
"""
from __future__ import annotations

import numpy as np
import torch

from ._lib import PlankB200Error, call


def _offsets(counts, device):
    off = np.zeros(len(counts) + 1, dtype=np.int32)
    np.cumsum(counts, out=off[1:])
    return torch.from_numpy(off).to(device)


def tokenize_lines(lines, views, types, data_cfg, token, device='cuda'):
    """lines: list (one per drawing) of [n_i, 4] float arrays in [-1, 1]; views / types: lists of [n_i] int arrays (types may
    be None).  -> dict with the reference's keys (`input_value`, `input_pos`, `input_coord`, `input_view`[, `input_type`],
    `input_mask`) as [B, MAX_INPUT_LENGTH - 1] CUDA tensors, plus `kv_len` [B] int32."""
    B, S = len(lines), data_cfg.MAX_INPUT_LENGTH - 1
    dev = torch.device(device)
    counts = [len(l) for l in lines]
    cat = torch.from_numpy(np.concatenate([np.asarray(l, dtype=np.float64).reshape(-1, 4) for l in lines])).to(dev)
    vw = torch.from_numpy(np.concatenate([np.asarray(v, dtype=np.int64) for v in views])).to(dev)
    ty = torch.from_numpy(np.concatenate([np.asarray(t, dtype=np.int64) for t in types])).to(dev) if types is not None else None
    off = _offsets(counts, dev)
    planes = {k: torch.empty(B, S, device=dev, dtype=torch.int64) for k in ('input_value', 'input_pos', 'input_coord', 'input_view')}
    if ty is not None:
        planes['input_type'] = torch.empty(B, S, device=dev, dtype=torch.int64)
    mask = torch.empty(B, S, device=dev, dtype=torch.uint8)
    kv_len = torch.empty(B, device=dev, dtype=torch.int32)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    call('pa_tokenize_lines', cat.data_ptr(), vw.data_ptr(), ty.data_ptr() if ty is not None else None, off.data_ptr(), B, S,
         data_cfg.NUM_BITS, token.END, token.PAD, planes['input_value'].data_ptr(), planes['input_pos'].data_ptr(),
         planes['input_coord'].data_ptr(), planes['input_view'].data_ptr(),
         planes['input_type'].data_ptr() if ty is not None else None, mask.data_ptr(), kv_len.data_ptr(), err.data_ptr(),
         torch.cuda.current_stream().cuda_stream)
    if max(counts) * 4 + 1 > S:                      # host-side check of what the kernel flags in `err` (no sync needed)
        raise PlankB200Error(f'a drawing has {max(counts)} lines: {max(counts) * 4 + 1} tokens do not fit MAX_INPUT_LENGTH - 1 = {S}')
    planes['input_mask'] = mask.view(torch.bool)
    planes['kv_len'] = kv_len
    return planes


def tokenize_planks(coords, attach, data_cfg, token, device='cuda'):
    """coords: list of flat float arrays (planks x 6, in [-1, 1]); attach: list of flat int arrays (-1 = not attached, else
    the index of the earlier output position).  -> `output_value`, `output_label`, `output_mask` [B, MAX_OUTPUT_LENGTH]."""
    B, T = len(coords), data_cfg.MAX_OUTPUT_LENGTH
    dev = torch.device(device)
    counts = [len(np.asarray(c).reshape(-1)) for c in coords]
    cat = torch.from_numpy(np.concatenate([np.asarray(c, dtype=np.float64).reshape(-1) for c in coords])).to(dev)
    att = torch.from_numpy(np.concatenate([np.asarray(a, dtype=np.int64).reshape(-1) for a in attach])).to(dev)
    off = _offsets(counts, dev)
    value = torch.empty(B, T, device=dev, dtype=torch.int64)
    label = torch.empty(B, T, device=dev, dtype=torch.int64)
    mask = torch.empty(B, T, device=dev, dtype=torch.uint8)
    out_len = torch.empty(B, device=dev, dtype=torch.int32)
    err = torch.zeros(1, device=dev, dtype=torch.int32)
    if max(counts) + 1 > T:
        raise PlankB200Error(f'{max(counts)} output tokens + END do not fit MAX_OUTPUT_LENGTH = {T}')
    call('pa_tokenize_planks', cat.data_ptr(), att.data_ptr(), off.data_ptr(), B, T, data_cfg.NUM_BITS, token.END, token.PAD,
         data_cfg.VOCAB_SIZE, value.data_ptr(), label.data_ptr(), mask.data_ptr(), out_len.data_ptr(), err.data_ptr(),
         torch.cuda.current_stream().cuda_stream)
    return {'output_value': value, 'output_label': label, 'output_mask': mask.view(torch.bool), 'out_len': out_len}
