"""The drop-in boundary as north_star states it: `trainer_complete.py` runs UNCHANGED against the new module.

The unmodified reference trainer (`oracle/_ref/trainer_complete.py`, staged by `oracle/build_ref.py`; it travels to the
GPU box git-ignored) is imported with `shim/` in front of it on sys.path, so its `from plankassembly.models import
build_model` resolves to the B200 module, and with tests/_shims standing in for pytorch_lightning / detectron2 /
torchmetrics / the dataset package (absent from this image).  The hooks Lightning would call are then driven by hand
under the conditions the reference's configs set: `detect_anomaly: True` (ref: configs/train_complete.yaml:16),
validation/test under `torch.inference_mode()`, checkpoints with `model.`-prefixed keys, `strategy: ddp` (torch DDP).

The CPU part (no GPU needed) checks that the shim resolves and that the trainer constructs on top of it.
"""
import importlib
import importlib.util
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, 'oracle', '_ref')
have_ref = os.path.exists(os.path.join(REF, 'trainer_complete.py'))
needs_ref = pytest.mark.skipif(not have_ref, reason='oracle/_ref not staged (python oracle/build_ref.py in the dev container)')

from plankassembly_b200 import synthetic as syn  # noqa: E402


def hparams(cfg, train='0:8', valid='100:104', test='100:104', batch=4):
    h = json.loads(json.dumps(dict(cfg)))            # plain nested dict, as jsonargparse hands it to Trainer(hparams)
    h.update(ROOT='unused', DATASETS_TRAIN=train, DATASETS_VALID=valid, DATASETS_TEST=test, BATCH_SIZE=batch, NUM_WORKERS=0)
    return h


@pytest.fixture
def trainer_module(monkeypatch):
    """import the reference trainer with: shim (our plankassembly.models) > tests/_shims (Lightning & co) > oracle/_ref."""
    for name in [m for m in sys.modules if m.split('.')[0] in ('plankassembly', 'trainer_complete', 'third_party', 'dataset',
                                                               'pytorch_lightning', 'detectron2', 'torchmetrics')]:
        monkeypatch.delitem(sys.modules, name)
    for p in (REF, os.path.join(ROOT, 'tests', '_shims'), os.path.join(ROOT, 'shim')):
        monkeypatch.syspath_prepend(p)
    importlib.invalidate_caches()
    return importlib.import_module('trainer_complete')


def reference_models():
    """The reference's own models.py under a private module name (for side-by-side numbers in the same process)."""
    spec = importlib.util.spec_from_file_location('_ref_plank_models', os.path.join(REF, 'plankassembly', 'models.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def to_dev(batch, dev):
    return {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}


@needs_ref
def test_shim_resolves_to_b200_module(trainer_module):
    import plankassembly.models as pm
    import plankassembly_b200.models as ours
    assert pm.build_model is ours.build_model and pm.PlankModel is ours.PlankModel
    t = trainer_module.Trainer(hparams(syn.tiny_cfg()))
    assert type(t.model) is ours.PlankModel
    assert all(k.startswith('model.') for k in t.state_dict() if not k.startswith('criterion'))
    opt = t.configure_optimizers()['optimizer']
    assert isinstance(opt, torch.optim.Adam) and sum(p.numel() for g in opt.param_groups for p in g['params']) == \
        sum(p.numel() for p in t.model.parameters())
    # the reference's model has the same parameter names and shapes (checkpoint compatibility, ref README.md:120-123)
    ref = reference_models().build_model(t.cfg)
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == {k: tuple(v.shape) for k, v in t.model.state_dict().items()}


@needs_ref
@pytest.mark.gpu
def test_trainer_hooks_run_unchanged(trainer_module, tmp_path):
    from _util import trained_tiny_state_dict
    dev = torch.device('cuda', 0)
    cfg = syn.tiny_cfg(dropout=0.1)
    torch.manual_seed(2022)
    t = trainer_module.Trainer(hparams(cfg)).to(dev)
    t.logger.log_dir = str(tmp_path)
    opt = t.configure_optimizers()['optimizer']

    # ---- fit: training_step under detect_anomaly (configs/train_complete.yaml:16), Adam from model.parameters()
    t.train()
    losses = []
    with torch.autograd.set_detect_anomaly(True):
        for i, batch in enumerate(t.train_dataloader()):
            loss = t.training_step(to_dev(batch, dev), i)
            assert loss.dim() == 0 and loss.requires_grad
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(loss.item())
    assert len(losses) == 2 and all(np.isfinite(losses))
    assert len(t.logged['train/loss']) == 2 and len(t.logged['train/accuracy']) == 2

    # ---- checkpoint round trip with Lightning's `model.` prefix, then the trained tiny weights for a meaningful validation
    sd = {k: v.clone() for k, v in t.state_dict().items()}
    t2 = trainer_module.Trainer(hparams(cfg)).to(dev)
    t2.load_state_dict(sd)
    assert all(torch.equal(a, b) for a, b in zip(t.model.state_dict().values(), t2.model.state_dict().values()))
    t.load_state_dict({'model.' + k: v for k, v in trained_tiny_state_dict().items()}, strict=False)

    # ---- validation + test under inference_mode, twice in a row (persistent decode buffers are reused), then train again
    ref_model = reference_models().build_model(t.cfg)
    ref_model.load_state_dict(trained_tiny_state_dict())
    ref_model.eval()
    t.eval()
    for rep in range(2):
        t.criterion.reset()
        with torch.inference_mode():
            for i, batch in enumerate(t.val_dataloader()):
                t.validation_step(to_dev(batch, dev), i)
            t.validation_epoch_end(None)
    with torch.inference_mode():
        t.criterion.reset()
        from plankassembly_b200 import postprocess
        ours_dir = tmp_path / 'ours'
        for i, batch in enumerate(t.test_dataloader()):
            t.test_step(to_dev(batch, dev), i)
            # the batched writer produces the same files as the reference's per-sample loop, byte for byte
            db = to_dev(batch, dev)
            o = t.model(db)
            postprocess.write_pred_jsons(str(ours_dir), postprocess.pred_json_records(
                batch['name'], o['samples'], o['attach'], db['output_value'], t.cfg.TOKEN.END, t.cfg.THRESHOLD))
            with torch.no_grad():
                ref_out = ref_model(batch)                       # the reference itself, on the host
            ours = t.model(to_dev(batch, dev))
            assert torch.equal(ours['samples'].cpu(), ref_out['samples']) and torch.equal(ours['attach'].cpu(), ref_out['attach'])
        t.test_epoch_end(None)
    assert t.logged['val/fmeasure'][0] == t.logged['val/fmeasure'][1]
    files = sorted(os.listdir(tmp_path / 'pred_jsons'))
    assert files == [f'synthetic_{i:05d}.json' for i in range(100, 104)]
    for f in files:
        assert open(tmp_path / 'pred_jsons' / f, 'rb').read() == open(tmp_path / 'ours' / 'pred_jsons' / f, 'rb').read(), f
    rec = json.load(open(tmp_path / 'pred_jsons' / files[0]))
    assert set(rec) == {'prediction', 'attach', 'groundtruth', 'precision', 'recall', 'fmeasure'}
    t.train()
    with torch.autograd.set_detect_anomaly(True):
        batch = next(iter(t.train_dataloader()))
        t.training_step(to_dev(batch, dev), 0).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in t.model.parameters())

    # ---- sideface batches carry no input_type (ref: trainer_sideface.py)
    batch = {k: v for k, v in to_dev(batch, dev).items() if k != 'input_type'}
    t.training_step(batch, 0).backward()


def _ddp_worker(rank, world, port, out):
    for p in (os.path.join(ROOT, 'shim'), ROOT, os.path.join(ROOT, 'tests')):
        sys.path.insert(0, p)
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dev = torch.device('cuda', rank)
    torch.cuda.set_device(dev)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    from plankassembly.models import build_model
    from plankassembly_b200.parallel import shard_indices
    cfg = syn.tiny_cfg()
    m = build_model(cfg)
    m.load_state_dict(syn.init_state_dict(cfg))
    ddp = DDP(m.to(dev).train(), device_ids=[rank])            # what Lightning's `strategy: ddp` builds
    batch = to_dev(syn.batch_for(cfg, shard_indices(8, rank, world)), dev)
    for _ in range(2):                                          # second pass: bucket views rebuilt, grads accumulate into them
        ddp.zero_grad()
        ddp(batch)['loss'].backward()
    torch.cuda.synchronize()
    if rank == 0:
        torch.save({n: p.grad.cpu() for n, p in m.named_parameters()}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.timeout(600)
@pytest.mark.parametrize('world', [1, 2])
def test_torch_ddp_gradients_match_single_process(world, tmp_path):
    """The module wrapped in torch DistributedDataParallel over NCCL (ref: configs/train_complete.yaml:18 `strategy: ddp`):
    gradients after DDP's bucketed all-reduce == mean over ranks of the single-process gradients of each shard."""
    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    import torch.multiprocessing as mp
    from plankassembly_b200.models import build_model
    from plankassembly_b200.parallel import shard_indices
    out = str(tmp_path / 'g.pt')
    mp.spawn(_ddp_worker, args=(world, 29650 + world, out), nprocs=world, join=True)
    got = torch.load(out)
    cfg = syn.tiny_cfg()
    m = build_model(cfg)
    m.load_state_dict(syn.init_state_dict(cfg))
    m = m.cuda().train()
    for r in range(world):
        (m(to_dev(syn.batch_for(cfg, shard_indices(8, r, world)), 'cuda'))['loss'] / world).backward()
    gn = float(torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters())))
    for n, p in m.named_parameters():
        assert (got[n] - p.grad.cpu()).abs().max().item() <= 1e-5 * gn + 1e-4 * p.grad.abs().max().item(), n
