"""batched, device-side post-processing of greedy-decode outputs for validation_step / test_step.

What the reference does per drawing in Python (ref: models.py:258-265 `parse_sequence` twice per sample; trainer_complete.py:
76-80 / 97-101 zero-extent filter; third_party/boxes.py:197-242 IoU matrix; third_party/matcher.py:28-61 Hungarian
assignment) becomes: ONE parse launch per token tensor, ONE filter + IoU launch per batch, one device->host copy, and the
Hungarian assignment on the host (scipy), as SURVEY 8(f) ranks it.  `eval_step` uses `parse_batch` for its `predicts` /
`groundtruths` lists; `batched_prf` is the B200-native replacement for the per-sample matcher loop and returns the very
numbers `HungarianMatcher.forward` returns (tests/test_gpu_postprocess.py checks it against the reference's own matcher).
"""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from ._lib import call


def _stream():
    return torch.cuda.current_stream().cuda_stream


def parse_batch(seq, end_token, dof=6):
    """seq [B, n] int64 (CUDA) -> planks [B, p_max, dof] int64, n_planks [B] int32, keep [B, p_max] uint8 (device tensors)."""
    assert seq.is_cuda and seq.dtype == torch.int64 and seq.dim() == 2
    if seq.stride(1) != 1:
        seq = seq.contiguous()
    B, n = seq.shape
    p_max = max(n // dof, 1)
    planks = torch.empty(B, p_max, dof, device=seq.device, dtype=torch.int64)
    n_planks = torch.empty(B, device=seq.device, dtype=torch.int32)
    keep = torch.empty(B, p_max, device=seq.device, dtype=torch.uint8)
    call('pa_parse_sequences', seq.data_ptr(), seq.stride(0), B, n, int(end_token), dof, planks.data_ptr(), p_max,
         n_planks.data_ptr(), keep.data_ptr(), _stream())
    return planks, n_planks, keep


def plank_lists(planks, n_planks):
    """-> list of [n_i, dof] int64 views (what ref models.py:309-315 collects), one host sync for the counts."""
    counts = n_planks.tolist()
    return [planks[i, :c] for i, c in enumerate(counts)]


def batched_iou(pred, pred_keep, n_pred, gt, n_gt):
    """IoU of every kept predicted plank (after plank 0) against every ground-truth plank (after plank 0), all drawings at once.
    -> iou [B, p_max-1, g_max-1] fp32, n_rows [B] int32, row_src [B, p_max-1] int32."""
    B, p_max, dof = pred.shape
    g_max = gt.shape[1]
    assert dof == 6 and gt.shape[2] == 6
    if p_max < 2 or g_max < 2:
        z = torch.zeros(B, max(p_max - 1, 0), max(g_max - 1, 0), device=pred.device)
        return z, torch.zeros(B, device=pred.device, dtype=torch.int32), torch.full((B, max(p_max - 1, 0)), -1, device=pred.device, dtype=torch.int32)
    iou = torch.empty(B, p_max - 1, g_max - 1, device=pred.device, dtype=torch.float32)
    n_rows = torch.empty(B, device=pred.device, dtype=torch.int32)
    row_src = torch.empty(B, p_max - 1, device=pred.device, dtype=torch.int32)
    call('pa_plank_iou', pred.data_ptr(), pred_keep.data_ptr(), n_pred.data_ptr(), p_max, gt.data_ptr(), n_gt.data_ptr(), g_max, B,
         iou.data_ptr(), n_rows.data_ptr(), row_src.data_ptr(), _stream())
    return iou, n_rows, row_src


def batched_prf(samples, output_value, end_token, threshold=0.5, dof=6):
    """Precision / recall / F1 of every drawing of a batch, as ref trainer_complete.py:97-104 + third_party/matcher.py score
    them: samples = decoded tokens [B, n], output_value = ground-truth tokens [B, T].  -> float64 array [B, 3]."""
    pred, n_pred, keep = parse_batch(samples, end_token, dof)
    gt, n_gt, _ = parse_batch(output_value, end_token, dof)
    return _prf_from(pred, n_pred, keep, gt, n_gt, threshold)


def _prf_from(pred, n_pred, keep, gt, n_gt, threshold):
    from scipy.optimize import linear_sum_assignment
    iou, n_rows, _ = batched_iou(pred, keep, n_pred, gt, n_gt)
    iou_h, rows_h, ng_h = iou.cpu().numpy(), n_rows.tolist(), n_gt.tolist()           # the one device->host copy
    out = np.zeros((len(rows_h), 3))
    for b, (nr, ng) in enumerate(zip(rows_h, ng_h)):
        ng = max(ng - 1, 0)
        if nr == 0 or ng == 0:
            continue
        m = iou_h[b, :nr, :ng]
        cost = np.full((nr, ng), 100000)
        cost[m > threshold] = -1
        ri, ci = linear_sum_assignment(cost)
        tp = np.float32((m[ri, ci] >= threshold).sum())
        prec, rec = tp / np.float32(nr), tp / np.float32(ng)                          # the reference divides fp32 tensors
        out[b] = (prec, rec, prec * rec * 2 / (prec + rec + np.float32(1e-10)))
    return out


def pred_json_records(names, samples, attach, output_value, end_token, threshold=0.5, dof=6):
    """the per-drawing records `test_step` dumps (ref trainer_complete.py:97-118) for a whole batch, from the
    batched device-side post-processing: {"prediction", "attach", "groundtruth", "precision", "recall", "fmeasure"} with the
    zero-extent planks dropped, `attach` cut to the kept prediction's length, metrics as Python floats of the fp32 values.
    -> list of (name, dict)."""
    pred, n_pred, keep = parse_batch(samples, end_token, dof)
    gt, n_gt, _ = parse_batch(output_value, end_token, dof)
    prf = _prf_from(pred, n_pred, keep, gt, n_gt, threshold)
    pred_h, keep_h, np_h = pred.cpu().numpy(), keep.cpu().numpy().astype(bool), n_pred.tolist()
    gt_h, ng_h, att_h = gt.cpu().numpy(), n_gt.tolist(), attach.cpu().numpy()
    out = []
    for b, name in enumerate(names):
        kept = pred_h[b, :np_h[b]][keep_h[b, :np_h[b]]]                               # plank 0 + non-degenerate planks
        att = att_h[b, :kept.size].reshape(-1, dof)                                   # ref: atta[:len(valid_pred.flatten())]
        out.append((name, {'prediction': kept.tolist(), 'attach': att.tolist(), 'groundtruth': gt_h[b, :ng_h[b]].tolist(),
                           'precision': float(prf[b, 0]), 'recall': float(prf[b, 1]), 'fmeasure': float(prf[b, 2])}))
    return out


def write_pred_jsons(log_dir, records):
    """Byte-compatible with the reference's writer (ref trainer_complete.py:110-118): one `<name>.json` per drawing under
    `<log_dir>/pred_jsons`, json.dump(indent=4, separators=(', ', ': '))."""
    d = os.path.join(log_dir, 'pred_jsons')
    os.makedirs(d, exist_ok=True)
    for name, rec in records:
        with open(os.path.join(d, f'{name}.json'), 'w') as f:
            json.dump(rec, f, indent=4, separators=(', ', ': '))
