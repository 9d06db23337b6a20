"""device-side parse_sequence + zero-extent filter + 3-D IoU (csrc/postprocess.cu) against the reference's own
code: `PlankModel.parse_sequence` semantics (ref models.py:258-265), the trainer's filter (ref trainer_complete.py:100-101)
and `third_party/matcher.py` / `boxes.py` imported unmodified from oracle/_ref when staged (else the restated metric of
tests/_util.py)."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from _util import plank_prf  # noqa: E402
from plankassembly_b200 import postprocess, synthetic as syn  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'oracle', '_ref')
END, PAD = syn.END, syn.PAD


def ref_parse(sequence, dof=6):
    valid = torch.cumsum(sequence == END, 0) == 0
    v = sequence[valid]
    n = len(v) // dof
    return v[:n * dof].reshape(-1, dof)


def random_sequences(B, n, seed):
    """Plank-like token rows with every edge case: END first, no END at all, END mid-plank, zero-extent planks, repeats after END."""
    g = np.random.default_rng(seed)
    seq = g.integers(0, 512, (B, n))
    for b in range(B):
        np_ = int(g.integers(0, n // 6 + 1))
        for j in range(np_):
            lo = g.integers(0, 400, 3)
            ext = g.integers(0, 100, 3) * (g.random(3) > 0.15)          # some zero extents
            seq[b, j * 6:j * 6 + 3], seq[b, j * 6 + 3:j * 6 + 6] = lo, lo + ext
        kind = b % 5
        if kind == 0:
            seq[b, 0] = END
        elif kind == 1:
            pass                                                        # no END anywhere
        elif kind == 2:
            seq[b, min(np_ * 6 + 2, n - 1)] = END                       # END in the middle of a plank
        else:
            seq[b, min(np_ * 6, n - 1)] = END
            if np_ * 6 + 9 < n:
                seq[b, np_ * 6 + 9] = END                               # tokens (and another END) after the first END
    return torch.from_numpy(seq).long()


@pytest.mark.parametrize('B,n', [(1, 5), (7, 64), (64, 128), (33, 256)])
def test_parse_batch_matches_parse_sequence(B, n):
    seq = random_sequences(B, n, seed=B * 1000 + n)
    planks, n_planks, keep = postprocess.parse_batch(seq.cuda(), END)
    lists = postprocess.plank_lists(planks, n_planks)
    for b in range(B):
        ref = ref_parse(seq[b])
        assert torch.equal(lists[b].cpu(), ref), b
        if len(ref) > 1:
            valid = torch.all(torch.abs(ref[1:, 3:] - ref[1:, :3]) != 0, dim=1)        # ref trainer_complete.py:100
            assert torch.equal(keep[b, 1:len(ref)].cpu().bool(), valid), b
        assert bool(keep[b, len(ref):].sum() == 0) and (len(ref) == 0 or bool(keep[b, 0] == 1))


def reference_matcher():
    if not os.path.exists(os.path.join(REF, 'third_party', 'matcher.py')):
        return None
    if REF not in sys.path:
        sys.path.insert(0, REF)
    from third_party.matcher import build_matcher
    return build_matcher(0.5)


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_batched_iou_and_prf_match_reference_matcher(seed):
    B, n = 24, 128
    gt = random_sequences(B, n, seed)
    pred = gt.clone()
    g = np.random.default_rng(seed + 100)
    jit = torch.from_numpy(g.integers(-3, 4, pred.shape)) * torch.from_numpy(g.random(pred.shape) < 0.3)
    pred = torch.where(pred < 512, (pred + jit).clamp(0, 511), pred)                    # perturbed predictions, END kept
    prf = postprocess.batched_prf(pred.cuda(), gt.cuda(), END, 0.5)
    matcher = reference_matcher()
    pl, npl, keep = postprocess.parse_batch(pred.cuda(), END)
    gl, ngl, _ = postprocess.parse_batch(gt.cuda(), END)
    iou, n_rows, row_src = postprocess.batched_iou(pl, keep, npl, gl, ngl)
    for b in range(B):
        p, t = ref_parse(pred[b]), ref_parse(gt[b])
        if len(p) == 0 or len(t) == 0:
            assert (prf[b] == 0).all()
            continue
        valid = torch.all(torch.abs(p[1:, 3:] - p[1:, :3]) != 0, dim=1)
        vp = p[1:][valid]
        assert int(n_rows[b]) == len(vp)
        assert row_src[b, :len(vp)].cpu().tolist() == (torch.nonzero(valid)[:, 0] + 1).tolist()
        if matcher is not None and len(vp) and len(t) > 1:
            from third_party.boxes import Boxes, pairwise_iou
            ref_iou = pairwise_iou(Boxes(vp), Boxes(t[1:]))
            assert torch.equal(iou[b, :len(vp), :len(t) - 1].cpu(), ref_iou), b                # bit-identical fp32
            ref_prf = [float(x) for x in matcher(vp, t[1:])]
            assert np.allclose(prf[b], ref_prf, rtol=0, atol=1e-7), (b, prf[b], ref_prf)
        assert np.allclose(prf[b], plank_prf(p, t, 0.5), rtol=0, atol=1e-6), b


def test_eval_step_lists_match_parse_sequence():
    from _util import trained_tiny_state_dict
    from plankassembly_b200.models import build_model
    cfg = syn.tiny_cfg()
    m = build_model(cfg)
    m.load_state_dict(trained_tiny_state_dict())
    m = m.cuda().eval()
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in syn.batch_for(cfg, range(10)).items()}
    out = m(batch)
    for i in range(10):
        assert torch.equal(out['predicts'][i], m.parse_sequence(out['samples'][i]))
        assert torch.equal(out['groundtruths'][i], m.parse_sequence(batch['output_value'][i]))
    prf = postprocess.batched_prf(out['samples'], batch['output_value'], cfg.TOKEN.END, cfg.THRESHOLD)
    ref = np.array([plank_prf(p, g, cfg.THRESHOLD) for p, g in zip(out['predicts'], out['groundtruths'])])
    assert np.allclose(prf, ref, atol=1e-6) and prf[:, 2].mean() > 0.9            # the overfit fixture decodes its own drawings
