"""A SHORT greedy decode for ncu launch lists: full model, BASELINE configs[2] shapes (S=512), batch B, but only N steps
(the launch list of one full 256-step decode at batch 1024 would hold ~230 k launches).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file out.csv \\
        python scripts/prof_decode.py 1024 8
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from plankassembly_b200 import synthetic as syn
from plankassembly_b200.models import build_model

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = syn.make_cfg(512, 8, 1024, 0.0, 6, 6, 513, steps, batch_size=B)          # max_output_length = steps
full = syn.config2(dropout=0.0)
sd = syn.init_state_dict(full)
model = build_model(cfg)
own = model.state_dict()
model.load_state_dict({k: (v if v.shape == own[k].shape else v[:own[k].shape[0]]) for k, v in sd.items()})   # shorter query_pos table
model = model.cuda().eval()
batch = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in syn.make_batch(range(B), 513, 256).items()}
with torch.no_grad():
    model(batch)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(batch)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print('profiled one decode:', B, 'sequences,', steps, 'steps')
