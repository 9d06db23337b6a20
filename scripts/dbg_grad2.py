"""Diagnostic: which reduced-precision component drives the gradient deviation? (GPU only)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from plankassembly_b200 import ops, synthetic as syn
from plankassembly_b200.models import build_model, PlankModel
from _util import case

name = sys.argv[1] if len(sys.argv) > 1 else 'tiny_trained'
cfg, sd, batch, g = case(name)
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}

def run(gemm, attn, bwd_tc, force_tf=None):
    ops.GEMM_IMPL = gemm
    ops.BWD_TC = bwd_tc
    m = build_model(cfg); m.load_state_dict(sd); m = m.cuda().train()
    m.attn_impl = attn
    if force_tf is not None:          # tensor-core attention with exact (cublas) projections
        m._impl = lambda: ops.ATTN_IMPL[attn]
    out = m.train_step(batch)
    out['loss'].backward()
    return out['loss'].item(), {n: p.grad.detach().double().clone() for n, p in m.named_parameters()}

ref_loss, ref = run('cublas', 'simt', False)
def report(tag, res):
    loss, gr = res
    num = sum(((gr[n] - ref[n]) ** 2).sum().item() for n in ref) ** 0.5
    den = sum((ref[n] ** 2).sum().item() for n in ref) ** 0.5
    worst = max(((gr[n] - ref[n]).norm() / (ref[n].norm() + 1e-30)).item() for n in ref)
    print(f'{tag:42s} loss {loss:.8f} (exact {ref_loss:.8f})  global rel-L2 {num / den:.3e}  worst param {worst:.3e}')
report('gemm tc   | attn simt', run('tc', 'simt', False))
report('gemm cublas | attn tc fwd + simt bwd', run('cublas', 'tc', False, True))
report('gemm cublas | attn tc fwd + tc bwd', run('cublas', 'tc', True, True))
report('gemm tc   | attn tc fwd + simt bwd', run('tc', 'tc', False))
report('gemm tc   | attn tc fwd + tc bwd', run('tc', 'tc', True))
