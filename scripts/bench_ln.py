"""Isolated timing of the residual + dropout + LayerNorm kernels on the model's shapes (GPU only): GB/s of algorithmic traffic."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plankassembly_b200 import ops
from plankassembly_b200._lib import call
d = 512
PD = float(os.environ.get('PDROP', 0.2))


def timeit(f, n=20):
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for rows in (32768, 16384):
    x, a = torch.randn(rows, d, device='cuda'), torch.randn(rows, d, device='cuda')
    ab, g, b = torch.randn(d, device='cuda'), torch.rand(d, device='cuda') + 0.5, torch.randn(d, device='cuda')
    y, yr = torch.empty_like(x), torch.empty_like(x)
    stats = torch.empty(rows, 2, device='cuda')
    s = torch.cuda.current_stream().cuda_stream
    fwd = lambda: call('pa_add_ln_fwd', x.data_ptr(), a.data_ptr(), ab.data_ptr(), g.data_ptr(), b.data_ptr(), 1.0, PD, 1, 2, rows, d,
                       y.data_ptr(), yr.data_ptr(), None, stats.data_ptr(), s)
    us = timeit(fwd)
    print(f'add_ln_fwd rows={rows}: {us:6.1f} us  {4 * rows * d * 4 / us / 1e3:6.0f} GB/s (x, a in; y, y_tf32 out)')
    dy, dy2, dx, da = torch.randn_like(x), torch.randn_like(x), torch.empty_like(x), torch.empty_like(x)
    dg, db, dab = torch.zeros(d, device='cuda'), torch.zeros(d, device='cuda'), torch.zeros(d, device='cuda')
    for two in (False, True):
        bwd = lambda: call('pa_add_ln_bwd', dy.data_ptr(), dy2.data_ptr() if two else None, yr.data_ptr(), stats.data_ptr(), g.data_ptr(), b.data_ptr(), PD,
                           1, 2, rows, d, dx.data_ptr(), da.data_ptr(), 1, dg.data_ptr(), db.data_ptr(), dab.data_ptr(), None, s)
        us = timeit(bwd)
        nt = 5 if two else 4
        print(f'add_ln_bwd rows={rows} dy2={two}: {us:6.1f} us  {nt * rows * d * 4 / us / 1e3:6.0f} GB/s ({nt - 2} in; dx, da out)')
    c = lambda: y.copy_(x)
    us = timeit(c)
    print(f'torch copy rows={rows}: {us:6.1f} us  {2 * rows * d * 4 / us / 1e3:6.0f} GB/s')
