"""Diagnostic: TF32 path vs exact fp32 path, per-parameter gradient agreement (GPU only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch
from plankassembly_b200 import ops, synthetic as syn
from plankassembly_b200.models import build_model
from _util import case

name = sys.argv[1] if len(sys.argv) > 1 else 'tiny_init'
cfg, sd, batch, g = case(name)
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
res = {}
for impl in ('cublas', 'tc'):
    ops.GEMM_IMPL = impl
    m = build_model(cfg); m.load_state_dict(sd); m = m.cuda().train()
    out = m.train_step(batch, return_dists=True)
    out['loss'].backward()
    res[impl] = (out['loss'].item(), out['dists'][0].detach().clone(), {n: p.grad.detach().clone() for n, p in m.named_parameters()})
print('loss', res['cublas'][0], res['tc'][0], 'golden', float(g['loss']))
d0, d1 = res['cublas'][1], res['tc'][1]
print('dists rel', ((d0 - d1).abs().max() / d0.abs().max()).item())
gn = dict(zip([str(n) for n in g['grad_names']], g['grad_norms']))
rows = []
for n in res['tc'][2]:
    a, b = res['cublas'][2][n].double(), res['tc'][2][n].double()
    rows.append((((a - b).norm() / (a.norm() + 1e-30)).item(), (b.norm() / (a.norm() + 1e-30)).item(), a.norm().item() / gn[n], n))
rows.sort(reverse=True)
print('rel_l2_err  norm_ratio(tc/exact)  exact/golden  name')
for r in rows[:25]:
    print(f'{r[0]:.3e}  {r[1]:.5f}  {r[2]:.5f}  {r[3]}')
print('...')
for r in rows[-5:]:
    print(f'{r[0]:.3e}  {r[1]:.5f}  {r[2]:.5f}  {r[3]}')
