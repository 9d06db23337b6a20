"""Isolated timing of the tensor-core attention kernels on the encoder shape (GPU only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plankassembly_b200 import ops
B, H, dh, L = 64, 8, 64, 512
d = H * dh
p = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
qkv = torch.randn(B, L, 3 * d, device='cuda').requires_grad_(True)
valid = int(sys.argv[2]) if len(sys.argv) > 2 else 400
kpm = torch.zeros(B, L, dtype=torch.uint8, device='cuda'); kpm[:, valid:] = 1
w = torch.randn(B, L, d, device='cuda')
def fwd():
    return ops.SelfAttention.apply(qkv, None, kpm, H, False, p, 1, True)
for _ in range(3):
    o = fwd(); o.backward(w)
torch.cuda.synchronize()
def timeit(f, n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
from plankassembly_b200 import _lib
ev = []
_lib.PROFILE_HOOK = {'pa_attn_fwd': (ev, None)}
for _ in range(10): o = fwd()
torch.cuda.synchronize()
_lib.PROFILE_HOOK = None
t_f = sum(a.elapsed_time(b) for a, b, _ in ev) / len(ev) * 1e3
ev = []
_lib.PROFILE_HOOK = {'pa_attn_bwd': (ev, None)}
for _ in range(5):
    o = fwd(); o.backward(w)
torch.cuda.synchronize()
_lib.PROFILE_HOOK = None
t_b = sum(a.elapsed_time(b) for a, b, _ in ev) / len(ev) * 1e3
fl = 4 * B * H * L * L * dh
print(f'p_drop={p} debug={os.environ.get("PLANK_B200_ATTN_DEBUG", "0")}: fwd {t_f:7.1f} us ({fl / t_f / 1e6:6.1f} TFLOP/s)   bwd(delta+dq+dkdv) {t_b:7.1f} us ({2.5 * fl / t_b / 1e6:6.1f} TFLOP/s)')
if int(os.environ.get('PLANK_B200_ATTN_DEBUG', '0')) & 64:
    import ctypes
    buf = (ctypes.c_ulonglong * 32)()
    fwd(); torch.cuda.synchronize()
    _lib.load().pa_debug_attn_prof(buf)
    v = list(buf)
    tot = v[31]
    print(f'CTA0 kernel cycles {tot}')
    print('producer waits: q_empty %5.1f%%  k_empty %5.1f%%  v_empty %5.1f%%' % tuple(100 * x / tot for x in v[0:3]))
    print('mma waits     : q_full %5.1f%%  k_full %5.1f%%  p_full %5.1f%%  o_empty %5.1f%%  v_full %5.1f%%' % tuple(100 * x / tot for x in v[8:13]))
    print('softmax waits : bar(bias) %5.1f%%  s_full %5.1f%%  bar(max) %5.1f%%  o_full %5.1f%%' % tuple(100 * x / tot for x in v[16:20]))
