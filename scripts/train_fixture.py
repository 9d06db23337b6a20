"""Train the full-size fixture model ON THE B200 through this repo's own train path and ship the weights back.

    gpurun -- python scripts/train_fixture.py --minutes 6 --memorise 32     (writes gpurun_out/fixture_weights_q8.npz + a log)

`--memorise N` (the recipe of the committed fixture): train ONLY on N configs[1]-shaped + N configs[3]-shaped drawings,
dropout 0, with the big matrices held ON the int8 storage grid during training (straight-through: the forward/backward
runs on the dequantised values, Adam updates fp32 masters), so the shipped int8 file reproduces the trained model exactly.
First attempt without it (profiles/r2_fixture_training.md): the eps=1 post-norm model grows its weights until it is so
high-gain that rounding a converged fp32 model to int8 afterwards takes the loss from 3.4 to 540.

TEST INFRASTRUCTURE.  The seeded-init fixtures have nearly flat output distributions
(top-1/top-2 margins down to 1e-6), so "bit-exact greedy tokens" on them is only checkable up to near-ties.  This script
produces a d=512, 6+6-layer model (tables sized for BOTH the configs[1] batches S=512/T=256 and the configs[3] batches
S=999/T=128: `synthetic.fixture_cfg`) whose distributions are peaked:

  set A   64 configs[1]-shaped drawings (idx 0..63, S=512, T=256)  -- memorised
  set B   64 configs[3]-shaped drawings (idx 0..63, S=999, T=128)  -- memorised
  pool    N configs[3]-shaped drawings (idx 1000..), a random half with line noise 0.05/0.10/0.20 -- for generalisation
  held-out configs[3]-shaped drawings idx 100..131: greedy-decode F1 is printed (must be well above 0 for configs[4])

The weights travel as int8 (groups of 32, fp16 scales: `synthetic.quantize_state_dict`); the DEQUANTISED values are the
fixture's weights, loaded alike by the reference (oracle/gen_golden.py, in the dev container) and by the CUDA path.
"""
from __future__ import annotations

import argparse
import copy
import math
import multiprocessing as mp
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from plankassembly_b200 import synthetic as syn  # noqa: E402


def _one(args):
    idx, mi, mo, noise = args
    return syn.make_sample(idx, mi, mo, noise_ratio=noise, canonical=True)


def collate(samples, dev):
    return {k: torch.from_numpy(np.stack([s[k] for s in samples])).to(dev) for k in samples[0]}


def take(big, idx):
    return {k: v[idx] for k, v in big.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--minutes', type=float, default=8.0)
    ap.add_argument('--pool', type=int, default=16384)
    ap.add_argument('--lr', type=float, default=3e-4)
    ap.add_argument('--dropout', type=float, default=0.1)
    ap.add_argument('--warm', type=int, default=300)
    ap.add_argument('--stop-loss', type=float, default=0.004)
    ap.add_argument('--memorise', type=int, default=0, help='train on this many drawings of each shape only')
    ap.add_argument('--store', default='fp16', choices=['fp16', 'q8'], help='fp16: matrices stored as fp16; q8: int8 groups of 32, trained on that grid (straight-through)')
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'fixture_weights_q8.npz'))
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    from _util import plank_prf
    from plankassembly_b200.models import build_model

    dev = torch.device('cuda', 0)
    cfg = syn.fixture_cfg(dropout=args.dropout)
    t0 = time.time()
    rng = np.random.default_rng(7)
    noise = rng.choice([0.0, 0.0, 0.0, 0.05, 0.10, 0.20], args.pool)
    jobs = ([(i, 513, 256, 0.0) for i in range(64)] + [(i, 1000, 128, 0.0) for i in range(64)]
            + [(i, 1000, 128, 0.0) for i in range(100, 132)]
            + [(1000 + i, 1000, 128, float(noise[i])) for i in range(args.pool)])
    with mp.Pool(min(32, os.cpu_count())) as pool:
        res = pool.map(_one, jobs, chunksize=64)
    set_a, set_b, held = collate(res[:64], dev), collate(res[64:128], dev), collate(res[128:160], dev)
    big = collate(res[160:], dev) if args.pool else None
    print(f'data: {len(res)} drawings in {time.time() - t0:.0f}s', flush=True)

    torch.manual_seed(2022)
    model = build_model(cfg)
    model.load_state_dict(syn.init_state_dict(cfg))
    model = model.to(dev).train()
    N = args.memorise
    if N:
        set_a, set_b = take(set_a, slice(0, N)), take(set_b, slice(0, N))
    QAT = bool(N) and args.store == 'q8'
    if QAT:
        big_params = [p for p in model.parameters() if p.dim() == 2 and p.shape[1] % syn.Q_GROUP == 0 and p.numel() >= 4096]
        masters = [p.detach().clone() for p in big_params]
        opt = torch.optim.Adam([{'params': masters}, {'params': [p for p in model.parameters() if all(p is not q for q in big_params)]}],
                               lr=args.lr, fused=True)

        def snap():
            """model weights <- int8-grid image of the fp32 masters (same arithmetic as synthetic.quantize_state_dict)"""
            with torch.no_grad():
                for p, m in zip(big_params, masters):
                    g = m.view(m.shape[0], -1, syn.Q_GROUP)
                    sc = (g.abs().amax(-1, keepdim=True) / 127.0).clamp_min(1e-12).half().float()
                    p.copy_((torch.round(g / sc).clamp_(-127, 127) * sc).view_as(m))
    else:
        opt = torch.optim.Adam(model.parameters(), lr=args.lr, fused=True)

    def tf_eval(batch):
        p, model.dropout = model.dropout, 0.0
        with torch.no_grad():
            out = model(batch)
        model.dropout = p
        return out['loss'].item(), out['accuracy'].item()

    def decode_f1(batch):
        model.eval()
        with torch.no_grad():
            out = model(batch)
        model.train()
        prf = np.array([plank_prf(p, g, cfg.THRESHOLD) for p, g in zip(out['predicts'], out['groundtruths'])])
        return prf.mean(0), out['samples'].shape[1]

    warm, budget = args.warm, args.minutes * 60
    ckpt, lr_scale = None, 1.0
    t_start, step, est_total = time.time(), 0, None
    while True:
        el = time.time() - t_start
        if el > budget:
            break
        # cosine over the time budget, linear warm-up over the first steps
        lr = lr_scale * args.lr * min(1.0, (step + 1) / warm) * (0.3 + 0.7 * 0.5 * (1 + math.cos(math.pi * min(1.0, el / budget))))
        for g in opt.param_groups:
            g['lr'] = lr
        if N:
            batch = set_a if step % 2 == 0 else set_b
            if QAT:
                snap()
                for m in masters:
                    m.grad = None
        elif step % 4 == 1:
            batch = set_a
        elif step % 8 == 3:
            batch = set_b
        else:
            batch = take(big, torch.randint(0, args.pool, (64,), device=dev))
        opt.zero_grad(set_to_none=True)
        out = model(batch)
        out['loss'].backward()
        if QAT:
            for p, m in zip(big_params, masters):
                m.grad = p.grad
        opt.step()
        if N and step % 100 == 0 and step > 0:
            if QAT:
                snap()
            (la, aa), (lb, ab) = tf_eval(set_a), tf_eval(set_b)
            if step % 500 != 0:
                print(f'step {step:6d} {el:5.0f}s A {la:.4f}/{aa:.4f} B {lb:.4f}/{ab:.4f}', flush=True)
            if aa >= 0.99999 and ab >= 0.99999 and max(la, lb) < args.stop_loss:
                print('memorised: stopping', flush=True)
                break
        if step % 500 == 0:
            la, aa = tf_eval(set_a)
            lb, ab = tf_eval(set_b)
            lh, ah = tf_eval(held)
            print(f'step {step:6d} {el:5.0f}s lr {lr:.2e} | pool loss {out["loss"].item():.4f} acc {out["accuracy"].item():.4f} | '
                  f'A {la:.4f}/{aa:.4f} B {lb:.4f}/{ab:.4f} held {lh:.4f}/{ah:.4f}', flush=True)
            if not (math.isfinite(out['loss'].item()) and math.isfinite(la)):
                if ckpt is None:
                    raise SystemExit('diverged before the first snapshot')
                print('  non-finite loss: restoring the last snapshot, halving the learning rate', flush=True)
                model.load_state_dict(ckpt[0]); opt.load_state_dict(ckpt[1]); lr_scale *= 0.5
                if QAT:
                    with torch.no_grad():
                        for m, c in zip(masters, ckpt[2]):
                            m.copy_(c)
            else:
                ckpt = ({k: v.clone() for k, v in model.state_dict().items()}, copy.deepcopy(opt.state_dict()),
                        [m.clone() for m in masters] if QAT else None)
        step += 1
    print(f'trained {step} steps in {time.time() - t_start:.0f}s', flush=True)
    if QAT:
        snap()
    print('held-out greedy decode P/R/F1, steps:', *decode_f1(held), flush=True)

    if args.store == 'q8':
        q = syn.quantize_state_dict(model.state_dict())
        np.savez(args.out, **q)
        stored = syn.dequantize_state_dict(q)
    else:
        args.out = args.out.replace('_q8', '_fp16')
        h = {k: v.detach().half().cpu() for k, v in model.state_dict().items()}
        np.savez_compressed(args.out, **{k: v.numpy() for k, v in h.items()})
        stored = {k: v.float() for k, v in h.items()}
    print(f'saved {args.out}: {os.path.getsize(args.out) / 2**20:.1f} MiB', flush=True)
    # what the fixture will actually hold: the stored (rounded) weights
    model.load_state_dict(stored)
    print('as stored: A %.4f/%.4f  B %.4f/%.4f  held %.4f/%.4f' % (*tf_eval(set_a), *tf_eval(set_b), *tf_eval(held)), flush=True)
    print('as stored, held-out greedy decode P/R/F1, steps:', *decode_f1(held), flush=True)
    if N:
        for ratio in (0.0, 0.05, 0.10, 0.20):
            nb = collate([syn.make_sample(i, 1000, 128, noise_ratio=ratio, canonical=True) for i in range(min(N, 16))], dev)
            print(f'noise {ratio}: set B[:16] greedy decode P/R/F1, steps:', *decode_f1(nb), flush=True)
    for name, bt in (('A', take(set_a, slice(0, 8))), ('B', take(set_b, slice(0, 8)))):
        model.eval()
        with torch.no_grad():
            o = model(bt)
        model.train()
        same = [bool((p.flatten() == g.flatten()).all()) if p.numel() == g.numel() else False for p, g in zip(o['predicts'], o['groundtruths'])]
        print(f'set {name}[:8] greedy decode == ground truth: {same}, steps {o["samples"].shape[1]}', flush=True)


if __name__ == '__main__':
    main()
