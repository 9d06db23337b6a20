"""Kernel time table via torch.profiler (CUPTI) for one train step or a few decode steps (GPU only)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from plankassembly_b200 import synthetic as syn
from plankassembly_b200.models import build_model

what = sys.argv[1] if len(sys.argv) > 1 else 'decode'
cfg = syn.config2(dropout=0.2)
m = build_model(cfg); m.load_state_dict(syn.init_state_dict(cfg)); m = m.cuda()
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in syn.batch_for(cfg, range(64)).items()}
if what == 'decode':
    os.environ['PLANK_B200_DECODE_GRAPH'] = '0'
    m.eval()
    from plankassembly_b200.decode import GreedyDecoder
    m._decoder_engine = GreedyDecoder(m)
    m._decoder_engine.use_graph = False
    with torch.no_grad():
        m(batch)
        m.max_output_length_saved = m.max_output_length
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            m(batch)
            torch.cuda.synchronize()
else:
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True)
    for _ in range(3):
        opt.zero_grad(set_to_none=True); m(batch)['loss'].backward(); opt.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        opt.zero_grad(set_to_none=True); m(batch)['loss'].backward(); opt.step()
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=22, max_name_column_width=70))
