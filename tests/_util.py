"""Shared helpers for the parity tests: golden fixtures + seeded weights/batches."""
import os

import numpy as np
import torch

from plankassembly_b200 import synthetic as syn

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

CASES = {
    'tiny_init': (syn.tiny_cfg, 'init'),
    'tiny_trained': (syn.tiny_cfg, 'trained'),
    'config1_init': (syn.config1, 'init'),
    'config2_init': (lambda: syn.config2(dropout=0.0), 'init'),
    'config4_init': (lambda: syn.config4(dropout=0.0), 'init'),      # train_visible.yaml shapes: S=999, T=128
}


FIXTURE_SHAPES = {'fixture_c2': (513, 256), 'fixture_c4': (1000, 128)}      # (MAX_INPUT_LENGTH, MAX_OUTPUT_LENGTH) of the batch


def fixture_state_dict():
    """Full-size trained fixture model (scripts/train_fixture.py; int8 groups, the dequantised values ARE the weights)."""
    return syn.dequantize_state_dict(np.load(os.path.join(GOLDEN, 'fixture_weights_q8.npz')))


def fixture_case(name):
    """-> (cfg, state_dict, batch, golden record) for the trained full-size fixture at the configs[1] / configs[3] shape."""
    g = golden(name)
    mi, mo = FIXTURE_SHAPES[name]
    batch = syn.make_batch([int(i) for i in g['indices']], mi, mo, canonical=True)
    return syn.fixture_cfg(dropout=0.0), fixture_state_dict(), batch, g


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + '.npz')))


def trained_tiny_state_dict():
    z = np.load(os.path.join(GOLDEN, 'tiny_trained_weights_fp16.npz'))
    return {k: torch.from_numpy(z[k]).float() for k in z.files}


def case(name):
    """-> (cfg, state_dict, batch, golden record) for a named fixture."""
    mk, kind = CASES[name]
    cfg = mk()
    sd = syn.init_state_dict(cfg) if kind == 'init' else trained_tiny_state_dict()
    g = golden(name)
    batch = syn.batch_for(cfg, [int(i) for i in g['indices']])
    return cfg, sd, batch, g


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs().max() / (b.abs().max() + 1e-30)).item()


def token_agreement(samples, attach, g, prefix='dec:', margin_floor=1e-4, strict=False):
    """Compare greedy tokens with the reference's.  Returns (ok, info); info is None when every token is identical.
    strict=True (trained fixtures): any difference fails.  Otherwise (seeded-init fixtures, whose distributions are nearly
    flat) a row may differ from the step on at which the reference's own relative top-1/top-2 margin is below
    `margin_floor` -- info then says how many rows took that escape, and the callers print it."""
    ref_s, ref_a, marg = g[prefix + 'samples'], g[prefix + 'attach'], g[prefix + 'margins']
    samples, attach = np.asarray(samples), np.asarray(attach)
    if samples.shape == ref_s.shape and (samples == ref_s).all() and (attach == ref_a).all():
        return True, None
    if strict:
        n = min(samples.shape[1], ref_s.shape[1])
        rows = [(b, int(np.nonzero((samples[b, :n] != ref_s[b, :n]) | (attach[b, :n] != ref_a[b, :n]))[0][0]))
                for b in range(ref_s.shape[0]) if ((samples[b, :n] != ref_s[b, :n]) | (attach[b, :n] != ref_a[b, :n])).any()]
        return False, {'strict': True, 'shapes': (samples.shape, ref_s.shape),
                       'first_mismatch (row, step, reference margin)': [(b, t, float(marg[b, t])) for b, t in rows]}
    bad = []
    n = min(samples.shape[1], ref_s.shape[1])
    for b in range(ref_s.shape[0]):
        diff = np.nonzero((samples[b, :n] != ref_s[b, :n]) | (attach[b, :n] != ref_a[b, :n]))[0]
        if len(diff):
            t = int(diff[0])
            bad.append((b, t, float(marg[b, t])))
    hard = [x for x in bad if x[2] > margin_floor]
    if samples.shape != ref_s.shape and not bad:
        hard.append(('length', samples.shape, ref_s.shape))
    return len(hard) == 0, {'near_tie_flips': bad, 'hard': hard, 'rows_escaped': len(bad) - len([x for x in hard if len(x) == 3 and not isinstance(x[0], str)])}


def plank_prf(pred, gt, threshold=0.5):
    """Precision / recall / F1 of one drawing as the reference scores it (TEST INFRASTRUCTURE, restated from
    ref: trainer_complete.py:100-103 -- drop zero-extent predictions, plank 0 (the bounding box) is excluded --,
    third_party/boxes.py:197-242 -- axis-aligned 3-D IoU, 0 where the boxes do not intersect -- and
    third_party/matcher.py:28-61 -- Hungarian assignment on cost -1 for IoU > threshold, 100000 otherwise; a matched pair
    counts when IoU >= threshold; F1 = 2PR / (P + R + 1e-10)).  pred, gt: integer [n, 6] (min xyz, max xyz)."""
    from scipy.optimize import linear_sum_assignment
    pred = np.asarray(torch.as_tensor(pred).cpu(), dtype=np.float64).reshape(-1, 6)
    gt = np.asarray(torch.as_tensor(gt).cpu(), dtype=np.float64).reshape(-1, 6)
    if len(pred):
        keep = np.all(np.abs(pred[1:, 3:] - pred[1:, :3]) != 0, axis=1)
        pred = np.concatenate([pred[:1], pred[1:][keep]])
    p, g = pred[1:], gt[1:]
    n_p, n_g = len(p), len(g)
    if n_p == 0 or n_g == 0:
        return 0.0, 0.0, 0.0
    lwh = np.minimum(p[:, None, 3:], g[None, :, 3:]) - np.maximum(p[:, None, :3], g[None, :, :3])
    inter = np.clip(lwh, 0, None).prod(-1)
    vol_p, vol_g = (p[:, 3:] - p[:, :3]).prod(-1), (g[:, 3:] - g[:, :3]).prod(-1)
    with np.errstate(divide='ignore', invalid='ignore'):
        iou = np.where(inter > 0, inter / (vol_p[:, None] + vol_g[None, :] - inter), 0.0)
    cost = np.full((n_p, n_g), 100000)
    cost[iou > threshold] = -1
    ri, ci = linear_sum_assignment(cost)
    tp = float((iou[ri, ci] >= threshold).sum())
    prec, rec = tp / n_p, tp / n_g
    return prec, rec, prec * rec * 2 / (prec + rec + 1e-10)
