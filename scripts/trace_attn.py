"""Event timeline of CTA 0 of the tensor-core attention forward (needs a library built with
PLANK_B200_NVCC_FLAGS=-DPA_ATTN_TRACE; run with PLANK_B200_ATTN_DEBUG=1024)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['PLANK_B200_ATTN_DEBUG'] = str(1024 | int(os.environ.get('PLANK_B200_ATTN_DEBUG', '0')))
import torch
from plankassembly_b200 import ops, _lib
B, H, dh, L = 64, 8, 64, 512
d = H * dh
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
qkv = torch.randn(B, L, 3 * d, device='cuda')
kpm = torch.zeros(B, L, dtype=torch.uint8, device='cuda'); kpm[:, 400:] = 1
with torch.no_grad():
    for _ in range(3):
        ops.SelfAttention.apply(qkv, None, kpm, H, False, p, 1, True)
torch.cuda.synchronize()
lib = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_ulonglong * (4 * 1024))()
n = (ctypes.c_int * 4)()
fn = getattr(lib, os.environ.get('TRACE_FN', 'pa_debug_attn_trace'))
assert fn(buf, n) == 0
names = {0: {0: 'q_empty ok', 2: 'k_empty ok', 3: 'v_empty ok'},
         1: {0: 'q_full ok', 1: 'k_full ok', 2: 'QK issued', 3: 'p_full ok', 4: 'o_empty ok', 5: 'v_full ok', 6: 'PV issued'},
         2: {0: 'tile start', 1: 'bar(bias)', 2: 's_full ok', 3: 'S loaded', 4: 'max done', 5: 'exp done', 6: 'P stored', 7: 'p_full arrive',
             8: 'o_full ok', 9: 'O accumulated', 10: 'item epilogue', 11: 'item end'}}
names[3] = names[2]
ev = []
for r in range(4):
    for i in range(n[r]):
        w = buf[r * 1024 + i]
        ev.append((w & ((1 << 56) - 1), r, w >> 56))
ev.sort()
t0 = ev[0][0]
lo, hi = int(os.environ.get('TRACE_FROM', 0)), int(os.environ.get('TRACE_TO', 30000))
role = ['TMA ', 'MMA ', 'SM-A', 'SM-B']
last = {}
for t, r, e in ev:
    if lo <= t - t0 <= hi:
        dt = t - last.get(r, t)
        print(f'{t - t0:8d}  {"            " * r}{role[r]} {names[r].get(e, e)} (+{dt})')
    last[r] = t
print('events per role:', list(n), ' span', ev[-1][0] - t0, 'cycles')
