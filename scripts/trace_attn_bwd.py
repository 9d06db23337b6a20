"""Event timeline of CTA 0 of the tensor-core attention dQ kernel (library built with PLANK_B200_NVCC_FLAGS=-DPA_ATTN_TRACE)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['PLANK_B200_ATTN_DEBUG'] = os.environ.get('TRACE_KERNEL', '1024')   # 1024 = dQ kernel, 2048 = dK/dV kernel
import torch
from plankassembly_b200 import ops, _lib
B, H, dh, L = 64, 8, 64, 512
d = H * dh
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
qkv = torch.randn(B, L, 3 * d, device='cuda').requires_grad_(True)
kpm = torch.zeros(B, L, dtype=torch.uint8, device='cuda'); kpm[:, 400:] = 1
w = torch.randn(B, L, d, device='cuda')
for _ in range(2):
    o = ops.SelfAttention.apply(qkv, None, kpm, H, False, p, 1, True); o.backward(w)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_ulonglong * (4 * 1024))()
n = (ctypes.c_int * 4)()
assert lib.pa_debug_attn_bwd_trace(buf, n) == 0
names = {0: {0: 'qdo_empty ok', 1: 'kv_empty ok'},
         1: {0: 'qdo_full+dq_empty ok', 1: 'kv_full ok', 2: 'S,dP issued', 3: 'ds_full ok', 4: 'dQ issued'},
         2: {0: 'tile start', 1: 'bar', 2: 'sdp_full ok', 3: 'S,dP loaded', 4: 'math done', 5: 'dS stored', 6: 'ds_full arrive', 7: 'dq_full ok', 8: 'item end', 9: 'item start', 10: 'qdo_full ok', 11: 'Q,dO in TMEM', 12: 'bias table written', 13: 'item bar', 14: 'acc loaded', 15: 'stores issued'}}
names[3] = names[2]
ev = []
for r in range(4):
    for i in range(n[r]):
        x = buf[r * 1024 + i]
        ev.append((x & ((1 << 56) - 1), r, x >> 56))
ev.sort()
t0 = ev[0][0]
lo, hi = int(os.environ.get('TRACE_FROM', 0)), int(os.environ.get('TRACE_TO', 30000))
role = ['TMA ', 'MMA ', 'EW-A', 'EW-B']
last = {}
for t, r, e in ev:
    if lo <= t - t0 <= hi:
        print(f'{t - t0:8d}  {"            " * r}{role[r]} {names[r].get(e, e)} (+{t - last.get(r, t)})')
    last[r] = t
print('events per role:', list(n), ' span', ev[-1][0] - t0, 'cycles')
