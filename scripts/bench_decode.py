"""Greedy-decode timing + token check of the decode engines on BASELINE configs[1]/[2] shapes (GPU only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plankassembly_b200 import synthetic as syn
from plankassembly_b200.models import build_model
from plankassembly_b200.decode import GreedyDecoder

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = syn.config2(dropout=0.0)
model = build_model(cfg)
model.load_state_dict(syn.init_state_dict(cfg))
model = model.cuda().eval()
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in syn.batch_for(cfg, range(B)).items()}
outs = {}
for mode in (sys.argv[2].split(',') if len(sys.argv) > 2 else ['graph', 'fused']):
    os.environ['PLANK_B200_DECODE'] = mode
    model._decoder_engine = GreedyDecoder(model)
    with torch.no_grad():
        o = model(batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        from plankassembly_b200 import _lib
        ev = []
        _lib.PROFILE_HOOK = {'pa_decode_fused': (ev, None)}
        e0.record()
        for _ in range(2):
            o = model(batch)
        e1.record(); torch.cuda.synchronize()
        _lib.PROFILE_HOOK = None
        if ev:
            print(f'   pa_decode_fused kernel alone: {sum(x.elapsed_time(y) for x, y, _ in ev) / len(ev):.2f} ms')
    ms = e0.elapsed_time(e1) / 2
    n = o['samples'].numel()
    outs[mode] = o
    print(f'{mode:6s} B={B}: {ms:8.2f} ms per decode, {o["samples"].shape[1]} steps, {n / ms * 1e3:10.0f} generated tok/s, {ms / o["samples"].shape[1] * 1e3:7.1f} us/step')
# prefill (encoder + cross-K/V projection) vs the token loop
with torch.no_grad():
    from plankassembly_b200 import ops
    def ev():
        e = torch.cuda.Event(enable_timing=True); e.record(); return e
    inputs = {k: v for k, v in batch.items() if k[:5] == 'input'}
    kpm = model._kpm(batch['input_mask'])
    for _ in range(2):
        e0 = ev(); ops.begin_step(); x = model._embed_input(inputs); memory, _ = model._encode(x, x, kpm); e1 = ev()
        model._decoder_engine.run(memory, kpm); e2 = ev()
    torch.cuda.synchronize()
    print(f'   encoder prefill {e0.elapsed_time(e1):.1f} ms, cross-K/V projection + token loop {e1.elapsed_time(e2):.1f} ms')
if len(outs) == 2:
    a, b = list(outs.values())
    print('samples equal:', torch.equal(a['samples'], b['samples']), ' attach equal:', torch.equal(a['attach'], b['attach']))

if os.environ.get('PLANK_B200_DECODE_PROF', '0') != '0':
    import ctypes
    from plankassembly_b200 import _lib
    buf = (ctypes.c_ulonglong * 16)()
    _lib.load().pa_debug_decode_prof(buf)
    names = ['gemm-qkv', 'self-attn', 'row', 'cross-attn', 'head', 'barriers', 'gemm-so/co', 'gemm-cq', 'gemm-f1', 'gemm-f2', 'gemm-heads']
    tot = sum(buf[:11])
    print('CTA0 phase profile (ms): ' + '  '.join(f'{n} {buf[i] / 1e6:.2f}' for i, n in enumerate(names)) + f'  total {tot / 1e6:.2f}')
    print(f'CTA0 GEMM items (all decodes since last read): {buf[15]}  staging {buf[12] / 1e6:.2f} ms  math {buf[13] / 1e6:.2f} ms  write {buf[14] / 1e6:.2f} ms')
