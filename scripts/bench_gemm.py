"""Isolated timing of the tensor-core GEMM on the model's shapes (GPU only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plankassembly_b200 import ops
shapes = [s for s in [('qkv  fwd', 32768, 1536, 512, {}), ('out  fwd', 32768, 512, 512, {}), ('ffn1 fwd', 32768, 1024, 512, {}), ('ffn2 fwd', 32768, 512, 1024, {}),
          ('qkv  dX ', 32768, 512, 1536, {'b_mn': True}), ('qkv  dW ', 1536, 512, 32768, {'a_mn': True, 'b_mn': True, 'split_k': 6, 'accumulate': True})] if len(sys.argv) < 2 or sys.argv[1] in s[0]]
modes = os.environ.get('PAIR_MODES', '2,1').split(',')
for name, M, N, K, kw in [(n + ' pair=' + m, M, N, K, dict(kw, _mode=m)) for (n, M, N, K, kw) in shapes for m in modes]:
    os.environ['PLANK_B200_GEMM_PAIR'] = kw.pop('_mode')
    if kw.get('a_mn'):
        a = torch.randn(K, M, device='cuda'); lda = M
    else:
        a = torch.randn(M, K, device='cuda'); lda = K
    if kw.get('b_mn'):
        b = torch.randn(K, N, device='cuda'); ldb = N
    else:
        b = torch.randn(N, K, device='cuda'); ldb = K
    c = torch.zeros(M, N, device='cuda')
    bias = None if kw else torch.randn(N, device='cuda')
    f = lambda: ops.gemm_tf32(a, b, c, M, N, K, lda=lda, ldb=ldb, ldc=N, bias=bias, **kw)
    for _ in range(3): f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f'{name} M={M} N={N} K={K}: {us:7.1f} us  {2 * M * N * K / us / 1e6:7.1f} TFLOP/s')

# yardstick: cuBLAS (torch.matmul) on the same shapes, TF32 allowed -- what the library reaches with fp32 operands in HBM
if os.environ.get('CUBLAS', '1') == '1':
    torch.backends.cuda.matmul.allow_tf32 = True
    for name, M, N, K in [('qkv  fwd', 32768, 1536, 512), ('out  fwd', 32768, 512, 512), ('ffn1 fwd', 32768, 1024, 512), ('ffn2 fwd', 32768, 512, 1024),
                          ('qkv  dX ', 32768, 512, 1536), ('qkv  dW ', 1536, 512, 32768)]:
        a = torch.randn(M, K, device='cuda'); b = torch.randn(N, K, device='cuda'); bias = torch.randn(N, device='cuda')
        f = (lambda: torch.addmm(bias, a, b.t())) if 'fwd' in name else (lambda: torch.matmul(a, b.t()))
        for _ in range(3): f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): f()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        print(f'cuBLAS tf32 {name} M={M} N={N} K={K}: {us:7.1f} us  {2 * M * N * K / us / 1e6:7.1f} TFLOP/s')
