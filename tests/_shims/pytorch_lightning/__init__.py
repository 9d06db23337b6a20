"""Stand-in for the slice of pytorch_lightning 1.7 that ref: trainer_complete.py touches (see ../README.md)."""
import types

import torch


class LightningModule(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.logged = {}
        self.logger = types.SimpleNamespace(log_dir='.')

    def save_hyperparameters(self, hparams):
        self.hparams = dict(hparams)

    def log(self, name, value, **kw):
        self.logged.setdefault(name, []).append(value.detach().float().cpu().item() if torch.is_tensor(value) else float(value))
