"""Host-side logic of the CUDA path that needs no GPU: the per-step zero block for gradient buffers and the kv_len
(1 + last non-PAD key) the tensor-core attention kernels use to skip key tiles that are PAD only."""
import torch

from plankassembly_b200 import ops


def test_kv_len_matches_definition():
    kpm = torch.zeros(5, 40, dtype=torch.uint8)
    kpm[0, 7:] = 1                      # 7 valid keys
    kpm[1, :] = 1; kpm[1, 0] = 0        # only key 0 valid
    kpm[2, 39] = 1                      # 39 valid
    kpm[3, 10:20] = 1                   # a hole in the middle does not shorten the row: last valid key is 39
    kpm[4, :] = 1                       # nothing valid: clamped to 1 (the kernels always visit one tile)
    out = ops._kv_len(kpm)
    assert out.dtype == torch.int32 and out.tolist() == [7, 1, 39, 40, 1]
    assert ops._kv_len(kpm) is out      # cached per mask tensor (one mask serves all layers of a step)
    kpm[0, 7] = 0                       # an in-place change bumps the version: recomputed
    assert ops._kv_len(kpm).tolist() == [8, 1, 39, 40, 1]
    assert ops._kv_len(None) is None
    with torch.inference_mode():        # inference tensors have no version counter: computed, not cached
        t = torch.zeros(2, 9, dtype=torch.uint8)
        t[1, 4:] = 1
        assert ops._kv_len(t).tolist() == [9, 4]


def test_zero_pool_learns_demand_and_hands_out_disjoint_zero_views():
    pool = ops._ZeroPool()
    dev = torch.device('cpu')
    pool.begin_step(dev)                # first step: nothing known yet -> plain torch.zeros
    a = pool.zeros((3, 5), dev); b = pool.zeros((70,), dev)
    assert a.shape == (3, 5) and b.shape == (70,) and float(a.abs().sum() + b.abs().sum()) == 0.0
    pool.begin_step(dev)                # second step: one block of the learned size (64-float granules)
    assert pool.buf is not None and pool.buf.numel() == 64 + 128
    a = pool.zeros((3, 5), dev); b = pool.zeros((70,), dev)
    a += 1.0
    assert float(b.abs().sum()) == 0.0 and a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr()
    assert a.data_ptr() % 256 == b.data_ptr() % 256         # 256-byte granules
    c = pool.zeros((1000,), dev)        # beyond the block: falls back, still zero
    assert float(c.abs().sum()) == 0.0 and c.untyped_storage().data_ptr() != a.untyped_storage().data_ptr()
    old = pool.buf
    pool.begin_step(dev)                # a new block per step: last step's views stay valid
    assert pool.buf is not old and float(a.sum()) == 15.0 and pool.buf.numel() >= 64 + 128 + 1024
