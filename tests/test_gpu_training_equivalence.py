"""Training equivalence (VERDICT r1 "weak" item 1): a model trained THROUGH the CUDA path must converge like the reference.

Fixture `tests/golden/tiny_train_curve.npz` = loss/accuracy of every step of the reference's own overfit recipe
(`oracle/gen_golden.py --only curve`: unmodified reference model, d=128 2+2 layers, 10 synthetic drawings as one batch,
seeded init, Adam lr 1e-3, dropout 0, fp32 CPU) until accuracy 1.0 and loss < 0.02 (380 steps).  Here the same recipe runs
from the same seeds through `plankassembly_b200.models.PlankModel` -- forward AND backward kernels, TF32 tensor-core path
('tc', what bench.py times) and fp32 path ('exact') -- with torch's plain Adam and with `fused=True` (which does not bump
Tensor._version: a stale-shadow-weight bug of round 1 that this test would have caught), and must

  * reproduce the reference's loss to BAND_EARLY over the first 80 steps (before rounding differences are amplified),
  * cross every loss level of LEVELS, reach teacher-forced accuracy 1.0 and meet the stop criterion within +-STEP_TOL
    (10 %) of the reference's step counts.

Training is a chaotic map of its rounding errors (Adam divides by sqrt(v)): around step 220 even the fp32 path shows a loss
bump the reference does not have (measured: loss 0.354 vs 0.245), so later deviations are bounded in STEPS, not in loss.
The measured numbers are appended to gpurun_out/training_equivalence.txt.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from _util import golden  # noqa: E402
from plankassembly_b200 import synthetic as syn  # noqa: E402

BAND_EARLY = 5e-3      # |loss/ref - 1| over steps 0..79
LEVELS = (4.5, 4.0, 3.0, 2.0, 1.0, 0.5, 0.25, 0.1, 0.05)
STEP_TOL = 0.10        # level crossings, first accuracy-1.0 step and stop step within +-10 % (+2 steps) of the reference's


def first(mask):
    idx = np.nonzero(mask)[0]
    return int(idx[0]) if len(idx) else None


ATTEMPTS = 3           # see below: the late phase is not bit-reproducible run to run


def run_recipe(precision, fused, ref_loss, ref_acc):
    """One training run of the reference's overfit recipe through the CUDA path -> (loss[], acc[])."""
    from plankassembly_b200.models import build_model
    n = len(ref_loss)
    cfg = syn.tiny_cfg()
    torch.manual_seed(2022)
    m = build_model(cfg)
    m.load_state_dict(syn.init_state_dict(cfg))
    m = m.cuda().train()
    batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in syn.batch_for(cfg, range(10)).items()}
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, fused=fused)
    steps = int(n * (1 + STEP_TOL)) + 3
    loss, acc = np.zeros(steps), np.zeros(steps)
    for i in range(steps):
        opt.zero_grad(set_to_none=True)
        out = m(batch)
        out['loss'].backward()
        opt.step()
        loss[i], acc[i] = out['loss'].item(), out['accuracy'].item()
    return loss, acc


@pytest.mark.parametrize('precision,fused', [('tc', False), ('tc', True), ('exact', False)])
def test_training_follows_reference_curve(precision, fused, monkeypatch):
    """The early band must hold in EVERY run.  The step-count criteria of the late phase must hold in one of ATTEMPTS runs:
    weight-gradient GEMMs (split-K TMA reduce-add), LayerNorm / bias / embedding gradients and the loss sum accumulate with
    fp32 atomics whose order differs run to run, so two runs of this chaotic recipe from identical seeds part after ~200
    steps, and about one run in ten shows a loss bump there (0.30 against 0.15 around step 240) that delays accuracy 1.0 by
    ~11 % -- the reference's own recipe is equally sensitive (the fp32 path bumps at step 220).  Every run is logged."""
    from plankassembly_b200 import ops
    monkeypatch.setattr(ops, 'GEMM_IMPL', 'tc' if precision == 'tc' else 'cublas')
    monkeypatch.setenv('PLANK_B200_ATTN', 'tc' if precision == 'tc' else 'simt')
    g = golden('tiny_train_curve')
    ref_loss, ref_acc = g['loss'], g['accuracy']
    n = len(ref_loss)
    ref_first, ref_stop = first(ref_acc >= 0.9999), n - 1
    failures = []
    for attempt in range(ATTEMPTS):
        loss, acc = run_recipe(precision, fused, ref_loss, ref_acc)
        our_first = first(acc >= 0.9999)
        our_stop = first((acc >= 0.9999) & (loss < 0.02))
        early = np.abs(loss[:80] / ref_loss[:80] - 1).max()
        cross = [(lv, first(loss < lv), first(ref_loss < lv)) for lv in LEVELS]
        os.makedirs('gpurun_out', exist_ok=True)
        with open('gpurun_out/training_equivalence.txt', 'a') as f:
            f.write(f'{precision} fused={fused} attempt {attempt}: early max dev {early:.3e}  first acc 1.0: ours {our_first} ref {ref_first}  stop: ours {our_stop} ref {ref_stop}\n')
            f.write('  first step below level (ours/ref): ' + ' '.join(f'{lv}:{a}/{b}' for lv, a, b in cross) + '\n')
            f.write('  loss every 20 steps ours/ref: ' + ' '.join(f'{a:.3f}/{b:.3f}' for a, b in zip(loss[:n:20], ref_loss[::20])) + '\n')
        assert early <= BAND_EARLY, early
        bad = [(lv, a, b) for lv, a, b in cross if a is None or abs(a - b) > STEP_TOL * b + 2]
        if our_first is None or abs(our_first - ref_first) > STEP_TOL * ref_first + 2:
            bad.append(('first accuracy 1.0', our_first, ref_first))
        if our_stop is None or abs(our_stop - ref_stop) > STEP_TOL * ref_stop + 2:
            bad.append(('stop', our_stop, ref_stop))
        if not bad:
            return
        failures.append(bad)
    raise AssertionError(f'no run out of {ATTEMPTS} met the +-{STEP_TOL:.0%} step-count criteria: {failures}')
