"""Ablation of the tensor-core GEMM on one shape (GPU only): epilogue forms x debug bits x grid size.
PLANK_B200_GEMM_DEBUG bits: 1 skip the epilogue body, 2 skip the stores, 4 skip the TMEM loads, 16 no proxy fence."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plankassembly_b200 import ops
M, N, K = int(os.environ.get('GM', 32768)), int(os.environ.get('GN', 1536)), int(os.environ.get('GK', 512))
a = torch.randn(M, K, device='cuda'); b = torch.randn(N, K, device='cuda'); c = torch.zeros(M, N, device='cuda'); bias = torch.randn(N, device='cuda')


def run(tag, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    f = lambda: ops.gemm_tf32(a, b, c, M, N, K, lda=K, ldb=K, ldc=N, bias=bias)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f'{tag:60s} {us:7.1f} us  {2 * M * N * K / us / 1e6:7.1f} TFLOP/s', flush=True)
    for k in env:
        os.environ.pop(k, None)


print(f'M={M} N={N} K={K}')
for epi in (0, 1, 2):
    for dbg in (0, 1, 2):
        run(f'EPI={epi} DEBUG={dbg}', PLANK_B200_GEMM_EPI=epi, PLANK_B200_GEMM_DEBUG=dbg)
for grid in ():
    run(f'EPI=0 mainloop only (DEBUG=1) grid={grid}', PLANK_B200_GEMM_EPI=0, PLANK_B200_GEMM_DEBUG=1, PLANK_B200_GEMM_GRID=grid)
    run(f'EPI=0 full grid={grid}', PLANK_B200_GEMM_EPI=0, PLANK_B200_GEMM_GRID=grid)
for pair in ():
    run(f'PAIR={pair} EPI=0 mainloop only', PLANK_B200_GEMM_PAIR=pair, PLANK_B200_GEMM_EPI=0, PLANK_B200_GEMM_DEBUG=1)
    run(f'PAIR={pair} EPI=0 full', PLANK_B200_GEMM_PAIR=pair, PLANK_B200_GEMM_EPI=0)
