"""N>1 host logic on CPU: world_size-2 gloo processes.  The CUDA kernels need a GPU, so the module
under the data-parallel wrapper is the oracle model; what is tested is OUR sharding + single flat
gradient all-reduce against single-process gradients with the reference's DDP semantics
(mean over ranks of per-rank mean losses)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    for p in (ROOT, os.path.join(ROOT, 'oracle')):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from plank_oracle import OraclePlankModel
    from plankassembly_b200 import synthetic as syn
    from plankassembly_b200.parallel import GradAllReduce, shard_indices
    torch.set_num_threads(2)
    cfg = syn.tiny_cfg()
    m = OraclePlankModel(cfg, syn.init_state_dict(cfg), requires_grad=True)
    m.training = True
    red = GradAllReduce(m.parameters())
    idx = shard_indices(8, rank, world)
    red.zero_grad()
    m.train_step(syn.batch_for(cfg, idx))['loss'].backward()
    red.sync()
    assert all(p.grad.data_ptr() >= red.flat.data_ptr() for p in m.parameters())   # gradients are views of the reduced buffer
    if rank == 0:
        torch.save({'flat': red.flat.clone(), 'idx': idx}, out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_flat_allreduce_matches_single_process(tmp_path):
    world, out = 2, str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(world, 29611, out), nprocs=world, join=True)
    got = torch.load(out)
    for p in (ROOT, os.path.join(ROOT, 'oracle')):
        sys.path.insert(0, p)
    from plank_oracle import OraclePlankModel
    from plankassembly_b200 import synthetic as syn
    from plankassembly_b200.parallel import shard_indices
    assert got['idx'] == [0, 2, 4, 6] and shard_indices(8, 1, 2) == [1, 3, 5, 7]
    cfg = syn.tiny_cfg()
    m = OraclePlankModel(cfg, syn.init_state_dict(cfg), requires_grad=True)
    m.training = True
    loss = sum(m.train_step(syn.batch_for(cfg, shard_indices(8, r, world)))['loss'] for r in range(world)) / world
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
    assert torch.allclose(got['flat'], ref, rtol=1e-4, atol=1e-7)


def test_shard_indices_cover_disjoint():
    from plankassembly_b200.parallel import shard_indices
    for world in (1, 2, 4, 8):
        shards = [shard_indices(64, r, world) for r in range(world)]
        flat = sorted(i for s in shards for i in s)
        assert flat == list(range(64)) and all(len(s) == 64 // world for s in shards)
