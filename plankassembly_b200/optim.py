"""Adam as ONE kernel over a flat fp32 master buffer that also refreshes the TF32 shadow weights.

The reference builds `torch.optim.Adam(self.model.parameters(), lr=cfg.LR)` (ref: trainer_complete.py:127-129; default
betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad) and that keeps working with this module.  `FusedAdam` is the
opt-in B200 form of the same update (same arithmetic, tests/test_gpu_optim.py holds it to 1e-6 against torch over 10 steps):

* at construction every parameter's storage is moved into one flat fp32 buffer (`p.data` becomes a view of it, values
  unchanged), with both moments and the TF32-rounded shadow copies as parallel flat buffers;
* `step()` is a single `pa_adam_flat` launch (csrc/adam.cu) instead of torch's ~8 multi_tensor_apply launches, and the
  shadow copies the tensor-core GEMMs read are written by the same pass instead of one `pa_round_tf32` launch per weight
  at the start of the next step (`ops.mark_shadow_fresh`).

`state_dict()` / `load_state_dict()` speak torch.optim.Adam's format (per-parameter `step`, `exp_avg`, `exp_avg_sq`), so a
checkpoint moves freely between the two optimizers.
"""
from __future__ import annotations

import math

import torch

from . import _lib, ops
from ._lib import call


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, shadow=True):
        defaults = dict(lr=lr, betas=betas, eps=eps)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError('FusedAdam keeps ONE flat buffer: a single parameter group')
        ps = [p for p in self.param_groups[0]['params'] if p.requires_grad]
        if not ps or any((not p.is_cuda) or p.dtype != torch.float32 for p in ps):
            raise _lib.PlankB200Error('FusedAdam needs fp32 CUDA parameters (no CPU fallback)')
        self._ps = ps
        dev = ps[0].device
        offs, n = [], 0
        for p in ps:
            offs.append(n)
            n += (p.numel() + 63) // 64 * 64                   # 256-byte granules: TMA / float4 alignment of every view
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.shadow = torch.zeros_like(self.flat) if shadow else None
        chunk = _lib.load().pa_adam_chunk_elems()
        c_off, c_par = [], []
        with torch.no_grad():
            for i, (p, o) in enumerate(zip(ps, offs)):
                view = self.flat[o:o + p.numel()].view_as(p)
                view.copy_(p)
                p.data = view                                   # the canonical parameter now lives in the flat buffer
                for s in range(0, p.numel(), chunk):
                    c_off.append(s)
                    c_par.append(i)
                self.state[p] = {'step': torch.zeros((), dtype=torch.float32),
                                 'exp_avg': self.exp_avg[o:o + p.numel()].view_as(p),
                                 'exp_avg_sq': self.exp_avg_sq[o:o + p.numel()].view_as(p)}
        self._offs = offs
        self._n_chunks = len(c_off)
        self._chunk_off = torch.tensor(c_off, dtype=torch.int64, device=dev)
        self._chunk_par = torch.tensor(c_par, dtype=torch.int32, device=dev)
        self._param_off = torch.tensor(offs, dtype=torch.int64, device=dev)
        self._param_len = torch.tensor([p.numel() for p in ps], dtype=torch.int64, device=dev)
        # per-step tables (gradient pointers, per-parameter bias corrections) travel through a RING of pinned host slots: the
        # host runs ahead of the GPU, so a slot may only be rewritten once the copy that read it has executed (its event)
        self._ring = 4
        self._gptr_host = [torch.zeros(len(ps), dtype=torch.int64).pin_memory() for _ in range(self._ring)]
        self._scal_host = [torch.zeros(len(ps), 2, dtype=torch.float32).pin_memory() for _ in range(self._ring)]
        self._slot_ev = [None] * self._ring
        self._gptr_dev = torch.zeros(len(ps), dtype=torch.int64, device=dev)
        self._scal_dev = torch.zeros(len(ps), 2, dtype=torch.float32, device=dev)
        self._steps = [0] * len(ps)
        self._t = 0
        self._keep = None

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self.param_groups[0]
        b1, b2 = g['betas']
        grads, ptrs = [], []
        for p in self._ps:
            gr = p.grad
            if gr is not None and (gr.dtype != torch.float32 or not gr.is_contiguous()):
                gr = gr.float().contiguous()
            grads.append(gr)
            ptrs.append(gr.data_ptr() if gr is not None else 0)
        if all(x is None for x in grads):
            return loss
        self._t += 1
        scal, cache = [], {}
        for i, (p, gr) in enumerate(zip(self._ps, grads)):
            if gr is not None:                                   # torch counts steps per parameter (a skipped one lags behind)
                self._steps[i] += 1
                self.state[p]['step'] += 1
            t = max(self._steps[i], 1)
            if t not in cache:
                cache[t] = (g['lr'] / (1 - b1 ** t), 1.0 / math.sqrt(1 - b2 ** t))
            scal.append(cache[t])
        slot = self._t % self._ring
        if self._slot_ev[slot] is not None:
            self._slot_ev[slot].synchronize()
        self._gptr_host[slot].copy_(torch.tensor(ptrs, dtype=torch.int64))
        self._scal_host[slot].copy_(torch.tensor(scal, dtype=torch.float32))
        self._gptr_dev.copy_(self._gptr_host[slot], non_blocking=True)
        self._scal_dev.copy_(self._scal_host[slot], non_blocking=True)
        self._slot_ev[slot] = torch.cuda.Event()
        self._slot_ev[slot].record()
        self._keep = grads                                       # gradients must outlive the launch
        call('pa_adam_flat', self.flat.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
             self.shadow.data_ptr() if self.shadow is not None else None, self._gptr_dev.data_ptr(), self._chunk_off.data_ptr(),
             self._chunk_par.data_ptr(), self._param_off.data_ptr(), self._param_len.data_ptr(), self._scal_dev.data_ptr(),
             self._n_chunks, b1, b2, g['eps'], torch.cuda.current_stream().cuda_stream)
        if self.shadow is not None:
            for p, o in zip(self._ps, self._offs):
                if p.dim() == 2:
                    ops.mark_shadow_fresh(p, self.shadow[o:o + p.numel()].view_as(p))
        return loss

    def load_state_dict(self, state_dict):
        """torch.optim.Adam format in; the loaded moments are copied INTO the flat buffers (views stay attached)."""
        views = {id(p): (self.state[p]['exp_avg'], self.state[p]['exp_avg_sq']) for p in self._ps}
        super().load_state_dict(state_dict)
        steps = []
        with torch.no_grad():
            for p in self._ps:
                st = self.state.get(p, {})
                m, v = views[id(p)]
                if 'exp_avg' in st and st['exp_avg'].data_ptr() != m.data_ptr():
                    m.copy_(st['exp_avg'])
                    v.copy_(st['exp_avg_sq'])
                st['exp_avg'], st['exp_avg_sq'] = m, v
                st['step'] = torch.as_tensor(float(st.get('step', 0.0)), dtype=torch.float32)
                steps.append(int(st['step']))
                self.state[p] = st
        self._steps = steps
        self._t = max(steps) if steps else 0
