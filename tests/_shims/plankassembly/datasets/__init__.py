"""plankassembly.datasets stand-in: LineDataset with the reference's constructor signature
(ref: plankassembly/datasets/line_data.py:16-32) producing synthetic drawings in its exact output layout."""
from torch.utils.data import Dataset

from plankassembly_b200 import synthetic as syn


class LineDataset(Dataset):
    def __init__(self, root, info_files, token, cfg, augmentation=False):
        self.indices = list(info_files)
        self.cfg = cfg
        self.augmentation = augmentation

    def __len__(self):
        return len(self.indices)

    def __getitem__(self, i):
        idx = self.indices[i]
        s = syn.make_sample(idx, self.cfg.MAX_INPUT_LENGTH, self.cfg.MAX_OUTPUT_LENGTH)
        return {'name': f'synthetic_{idx:05d}', **s}
