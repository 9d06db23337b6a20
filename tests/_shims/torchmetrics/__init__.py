"""torchmetrics.Metric stand-in for ref: plankassembly/metric.py (running sums, no distributed sync)."""
import torch


class Metric(torch.nn.Module):
    def __init__(self, **kw):
        super().__init__()
        self._defaults = {}

    def add_state(self, name, default, dist_reduce_fx=None):
        self._defaults[name] = default.clone()
        self.register_buffer(name, default.clone(), persistent=False)      # states follow .to(device) like torchmetrics'

    def reset(self):
        for k, v in self._defaults.items():
            setattr(self, k, v.clone().to(getattr(self, k).device))
