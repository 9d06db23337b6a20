"""Event timeline of CTA 0 of the tensor-core GEMM (library built with PLANK_B200_NVCC_FLAGS=-DPA_GEMM_TRACE)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['PLANK_B200_GEMM_DEBUG'] = str(8 | int(os.environ.get('EXTRA_DEBUG', '0')))
import torch
from plankassembly_b200 import ops, _lib
M, N, K = 32768, int(os.environ.get('GN', 1536)), int(os.environ.get('GK', 512))
a = torch.randn(M, K, device='cuda'); b = torch.randn(N, K, device='cuda'); c = torch.zeros(M, N, device='cuda'); bias = torch.randn(N, device='cuda')
for _ in range(2):
    ops.gemm_tf32(a, b, c, M, N, K, lda=K, ldb=K, ldc=N, bias=bias)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_ulonglong * (3 * 2048))(); n = (ctypes.c_int * 3)()
assert lib.pa_debug_gemm_trace(buf, n) == 0
names = {0: {0: 'empty ok (load issued next)'}, 1: {0: 'tmem_empty ok', 1: 'full ok', 2: 'k-block issued'},
         2: {0: 'epi tile start', 1: 'tmem_full ok', 2: 'store slot free', 3: 'epi tile end'}}
ev = []
for r in range(3):
    for i in range(n[r]):
        x = buf[r * 2048 + i]; ev.append((x & ((1 << 56) - 1), r, x >> 56))
ev.sort(); t0 = ev[0][0]
lo, hi = int(os.environ.get('TRACE_FROM', 0)), int(os.environ.get('TRACE_TO', 40000))
role = ['TMA', 'MMA', 'EPI']; last = {}
for t, r, e in ev:
    if lo <= t - t0 <= hi: print(f'{t - t0:8d}  {"                " * r}{role[r]} {names[r].get(e, e)} (+{t - last.get(r, t)})')
    last[r] = t
print('events', list(n), 'span', ev[-1][0] - t0)
