"""SM clock and board power while train steps run back to back (GPU only): is the step power-capped?"""
import os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pynvml
from plankassembly_b200 import synthetic as syn
from plankassembly_b200.models import build_model
from plankassembly_b200.optim import FusedAdam
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
B = int(os.environ.get('BATCH', 64)); SECS = float(os.environ.get('SECS', 4.0))
cfg = getattr(syn, os.environ.get('CFG', 'config2'))(dropout=0.2)
dev = torch.device('cuda', 0)
torch.manual_seed(2022)
model = build_model(cfg); model.load_state_dict(syn.init_state_dict(cfg)); model = model.to(dev).train()
opt = FusedAdam(model.parameters(), lr=cfg.LR)
batches = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.batch_for(cfg, range(i * B, (i + 1) * B)).items()} for i in range(2)]


def step(b):
    opt.zero_grad(set_to_none=True)
    out = model(b); out['loss'].backward(); opt.step()


for i in range(5): step(batches[i % 2])
torch.cuda.synchronize()
stop, out = threading.Event(), []


def sample():
    while not stop.is_set():
        out.append((time.time(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
        time.sleep(0.01)


th = threading.Thread(target=sample); th.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 0; t0 = time.time(); e0.record()
while time.time() - t0 < SECS:
    step(batches[n % 2]); n += 1
e1.record(); torch.cuda.synchronize(); stop.set(); th.join()
ms = e0.elapsed_time(e1) / n
clk = sorted(x[1] for x in out); pw = sorted(x[2] for x in out)
q = lambda a, f: a[min(len(a) - 1, int(f * len(a)))]
print(f'batch {B}: {ms:.2f} ms/step over {n} steps; SM MHz p10/p50/p90 = {q(clk, .1)}/{q(clk, .5)}/{q(clk, .9)}; '
      f'power W p10/p50/p90/max = {q(pw, .1):.0f}/{q(pw, .5):.0f}/{q(pw, .9):.0f}/{pw[-1]:.0f}; enforced limit {pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1e3:.0f} W')
