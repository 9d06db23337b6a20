"""Stage the UNMODIFIED reference files of the hot path under oracle/_ref/ so they travel to the GPU box.

TEST / BASELINE INFRASTRUCTURE.  `/root/reference` exists only in the dev container; `oracle/_ref/` is git-ignored (the
reference's sources never enter this repo's history) but NOT gpurun-ignored, so what this recipe stages there rides
along with the snapshot, exactly like the built `.so`.  The reference is pure Python over torch, so "building" it is a
verbatim file copy (checked by digest); nothing under `plankassembly_b200/` imports it.  Users:

  * `bench.py --impl reference` and `cpu_baseline` time `oracle/_ref/plankassembly/models.py` (the reference itself,
    `kind: "reference"`) on the box's host cores, falling back to the oracle port when the directory is absent;
  * `tests/test_trainer_boundary.py` drives the unmodified `trainer_complete.Trainer` (with stand-ins for Lightning,
    detectron2, torchmetrics and the dataset package, tests/_shims) against `shim/plankassembly/models.py`.

    python oracle/build_ref.py        (run by __graft_entry__.build() when /root/reference is present)
"""
from __future__ import annotations

import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('PLANK_REFERENCE', '/root/reference')
DST = os.path.join(HERE, '_ref')
FILES = ['plankassembly/models.py', 'plankassembly/metric.py', 'third_party/matcher.py', 'third_party/boxes.py',
         'trainer_complete.py', 'trainer_visible.py', 'configs/train_complete.yaml', 'configs/train_visible.yaml', 'LICENSE']


def _sha(p):
    with open(p, 'rb') as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(verbose=True):
    """-> True when oracle/_ref holds the reference files (copied now or already there)."""
    if not os.path.isdir(REF):
        ok = all(os.path.exists(os.path.join(DST, f)) for f in FILES)
        if verbose:
            print(f'oracle/_ref: {REF} not present; staged copy {"found" if ok else "MISSING"}')
        return ok
    manifest = []
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or _sha(dst) != _sha(src):
            shutil.copyfile(src, dst)
        manifest.append(f'{_sha(dst)}  {f}')
    with open(os.path.join(DST, 'MANIFEST.sha256'), 'w') as m:
        m.write('\n'.join(manifest) + '\n')
    if verbose:
        print(f'oracle/_ref: staged {len(FILES)} reference files from {REF}')
    return True


def available():
    return os.path.exists(os.path.join(DST, 'plankassembly', 'models.py'))


if __name__ == '__main__':
    build()
