"""bench.py -- shape-program tokens/sec of the PlankAssembly hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload config2|config4] [--impl b200|reference|torch_cuda]
                  [--decode-drawings D] [--decode-batch B] [--no-decode] [--no-cpu-baseline] [--no-torch-cuda]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one synthetic batch:
  train  (headline)  full model d=512 6+6 layers, per-GPU batch 64, configured dropout 0.2, forward + backward + gradient
         all-reduce (N>1) + Adam.  tokens = decoder positions B*T (the "shape program").
         --workload config2 (default) = BASELINE configs[1]: S=512 encoder / T=256 decoder positions;
         --workload config4           = BASELINE configs[3]: train_visible.yaml shapes S=999 / T=128 (per-rank batch 64,
                                        i.e. global 512 on 8 GPUs).
  decode (nested under "decode", BASELINE configs[2]): KV-cached greedy decode of D drawings per GPU (default 1000, rounded
         up to whole batches), max_len 256; tokens = generated positions.
Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = the same metric through the public PlankModel
API with pinned HOST batches copied in and the loss read back every step.  `roofline` = the dominant kernel FAMILY of the
step (the TF32 tensor-core GEMMs), measured live in a separate event-bracketed pass; `roofline_attention` the second one.
`cpu_baseline` / `--impl reference`: the UNMODIFIED reference model (oracle/_ref/plankassembly/models.py, staged by
oracle/build_ref.py; kind "reference") on the box's host cores, the oracle port (kind "port") when that directory is
absent.  `torch_cuda`: the same unmodified reference model on the same B200 through torch's own CUDA kernels
(cuBLAS + SDPA), with and without TF32 -- the second, tougher bar.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'shape-program tokens/sec (train step, fwd+bwd+Adam)'
UNIT = 'tokens/s'
DTYPE = 'tf32 (fp32 accumulate; fp32 master weights, fp32 residual stream)'

WORKLOADS = {
    'config2': ('config2', 'BASELINE configs[1]: train_complete full model d=512 H=8 ff=1024 6+6 layers, dropout 0.2'),
    'config4': ('config4', 'BASELINE configs[3]: train_visible.yaml shapes, full model d=512 H=8 ff=1024 6+6 layers, dropout 0.2'),
}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tf_burst': d['bf16_tflops'], 'tf_sustained': d['bf16_tflops_sustained'], 'src': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'src': 'fallback (B200_PROFILING.md)'}


def cpu_model():
    try:
        for line in open('/proc/cpuinfo'):
            if line.startswith('model name'):
                return line.split(':', 1)[1].strip()
    except OSError:
        pass
    return 'unknown'


def reference_models():
    """The unmodified reference module (oracle/_ref, git-ignored, staged by oracle/build_ref.py) or None."""
    path = os.path.join(ROOT, 'oracle', '_ref', 'plankassembly', 'models.py')
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location('_ref_plank_models', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.proc = index, None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                                          '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.summary = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if sm:
            self.summary = {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}


# ------------------------------------------------------------------------------------------------
def workload_cfg(name, dropout=0.2):
    from plankassembly_b200 import synthetic as syn
    return getattr(syn, WORKLOADS[name][0])(dropout=dropout)


def workload_config(name, B, S, T, world, extra=None):
    c = {'workload': f'{WORKLOADS[name][1]}, per-GPU batch {B}, S={S} encoder / T={T} decoder positions, fwd+bwd+allreduce+Adam',
         'global_batch': B * world, 'parallelism': f'dp{world}'}
    c.update(extra or {})
    return c


def step_flops(cfg, B, S, T):
    """Algorithmic FLOPs of one train step (fwd x3)."""
    d, ff, V = cfg.MODEL.NUM_MODEL, cfg.MODEL.NUM_FEEDFORWARD, cfg.DATA.VOCAB_SIZE
    Le, Ld = cfg.MODEL.NUM_ENCODER_LAYERS, cfg.MODEL.NUM_DECODER_LAYERS
    enc = Le * S * (8 * d * d + 4 * d * ff + 4 * S * d)
    dec = Ld * T * (12 * d * d + 4 * d * ff + 4 * T * d + 4 * S * d) + Ld * S * 4 * d * d
    heads = T * (2 * d * V + 2 * d * d + 2 * T * d + 2 * d)
    return 3.0 * B * (enc + dec + heads)


def _gemm_work(args):
    g = args[0]._obj
    return 2.0 * g.M * g.N * g.K * max(1, g.batch)


def _attn_work(mult):
    def f(args):
        a = args[0]._obj
        return mult * 4.0 * a.B * a.H * a.Lq * a.Lk * a.dh
    return f


def run_train(args, rank, world, local_rank):
    import torch.distributed as dist
    from plankassembly_b200 import _lib, synthetic as syn
    from plankassembly_b200.models import build_model
    from plankassembly_b200.parallel import GradAllReduce

    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if not _lib.load().pa_device_ok():
        raise SystemExit('bench.py needs a B200 (sm_100a): ' + _lib.load().pa_last_error().decode())
    cfg = workload_cfg(args.workload)
    B, S, T = args.batch, cfg.DATA.MAX_INPUT_LENGTH - 1, cfg.DATA.MAX_OUTPUT_LENGTH
    torch.manual_seed(2022)
    model = build_model(cfg)
    model.load_state_dict(syn.init_state_dict(cfg))
    model = model.to(dev).train()
    # N > 1: ONE NCCL all-reduce (average) per step over the concatenated gradients (parallel.py)
    reducer = GradAllReduce(model.parameters()) if world > 1 else None
    # Adam as one launch over a flat master buffer that also writes the TF32 shadow weights (optim.py);
    # --torch-adam times the reference's own optimizer object (torch.optim.Adam, fused) instead
    from plankassembly_b200.optim import FusedAdam
    opt = torch.optim.Adam(model.parameters(), lr=cfg.LR, fused=True) if args.torch_adam else FusedAdam(model.parameters(), lr=cfg.LR)

    # distinct synthetic drawings per rank (weak scaling: per-GPU batch fixed)
    n_host = 4
    host = [syn.batch_for(cfg, range((rank * n_host + i) * B, (rank * n_host + i + 1) * B)) for i in range(n_host)]
    for hb in host:
        for k, v in hb.items():
            if torch.is_tensor(v):
                hb[k] = v.pin_memory()
    resident = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in hb.items()} for hb in host]
    h2d = sum(v.numel() * v.element_size() for v in host[0].values() if torch.is_tensor(v))

    def step(batch):
        if reducer is not None:
            reducer.zero_grad()
        else:
            opt.zero_grad(set_to_none=True)
        out = model(batch)
        out['loss'].backward()
        if reducer is not None:
            reducer.sync()
        opt.step()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for i in range(args.warmup):
        step(resident[i % n_host])
    if args.profile_decode:
        model.load_state_dict(syn.init_state_dict(cfg))
        model.eval()
        with torch.no_grad():
            model(resident[0])
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            model(resident[0])
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        return {'profiled': 'one decode'}
    if args.profile_step:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step(resident[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return {'profiled': 'one step'}
    l0 = _lib.launch_count()
    # --- headline: device-resident timing, nothing else inside the timed region
    with ClockSampler(local_rank) as clk:
        ms = timed(lambda i: step(resident[i % n_host]), args.steps)
    launches = _lib.launch_count() - l0

    # --- end to end through the public API: pinned host batch -> device, result (loss, accuracy) copied back to pinned host
    # memory EVERY step.  The host consumes step i's numbers while step i+1 is already queued (the copy lands in a pinned ring
    # slot guarded by an event), as a training loop that logs its loss does; the last step's result is awaited inside the
    # timed region.
    ring = [torch.empty(2, dtype=torch.float32).pin_memory() for _ in range(2)]
    ring_ev = [torch.cuda.Event() for _ in range(2)]
    seen = []

    def e2e_step(i):
        hb = host[i % n_host]
        batch = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in hb.items()}
        out = step(batch)
        slot = i % 2
        ring[slot].copy_(torch.stack([out['loss'].detach(), out['accuracy'].detach()]), non_blocking=True)
        ring_ev[slot].record()
        if i > 0:                                      # read the PREVIOUS step's result (its copy has certainly landed or is awaited here)
            ring_ev[1 - slot].synchronize()
            seen.append(ring[1 - slot].tolist())
        if i == args.steps - 1:                        # ... and the last one before the clock stops
            ring_ev[slot].synchronize()
            seen.append(ring[slot].tolist())

    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    seen.clear()
    ms_e2e = timed(e2e_step, args.steps)
    assert len(seen) == args.steps and all(x[0] == x[0] for x in seen), 'e2e: every step must deliver a finite loss to the host'

    # --- roofline pass, OUTSIDE the headline timing: every GEMM / attention entry point bracketed with CUDA events on the
    # launching stream for a few steps; work per launch comes from the argument block of each call
    rec = {'pa_gemm_tf32': ([], _gemm_work), 'pa_attn_fwd': ([], _attn_work(1.0)), 'pa_attn_bwd': ([], _attn_work(2.5))}
    prof_steps = 3
    _lib.PROFILE_HOOK = rec
    for i in range(prof_steps):
        step(resident[i % n_host])
    torch.cuda.synchronize()
    _lib.PROFILE_HOOK = None

    def family(names):
        ev = [x for n in names for x in rec[n][0]]
        t_ms = sum(a.elapsed_time(b) for a, b, _ in ev)
        fl = sum(w for _, _, w in ev)
        return len(ev), t_ms, fl

    tokens = B * T * world
    ms_step = ms / args.steps
    pk = peaks()
    res = {
        'metric': METRIC, 'value': tokens * args.steps / (ms / 1e3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': DTYPE, 'data': 'synthetic',
        'config': workload_config(args.workload, B, S, T, world, {
            'attention': model.attn_impl, 'optimizer': 'torch.optim.Adam(fused=True)' if args.torch_adam else 'plankassembly_b200.optim.FusedAdam',
            'l2': 'per-step activations (>3 GB) exceed the 126 MB L2; no explicit flush',
            'e2e': 'pinned H2D of every batch + D2H of (loss, accuracy) every step; the host reads step i while step i+1 runs',
            'encoder_tokens_per_step': B * S * world}),
        'clocks': clk.summary,
        'e2e': {'value': tokens * args.steps / (ms_e2e / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 8},
        'gpu_launches': launches,
        'step_tflops': {'algorithmic_tflop_per_step': step_flops(cfg, B, S, T) / 1e12,
                        'achieved': step_flops(cfg, B, S, T) / (ms_step / 1e3) / 1e12,
                        'frac_of_bf16_sustained': step_flops(cfg, B, S, T) / (ms_step / 1e3) / 1e12 / pk['tf_sustained']},
    }
    # TF32 kind::tf32 MMAs run at half the bf16 rate: `peak` stays the measured bf16 number the contract names, `peak_tf32`
    # (= peak / 2) is the ceiling a TF32 kernel can actually reach.
    # `traffic`: dram__bytes_read + dram__bytes_write per launch of the family, OFFLINE from the ncu pass over one step of this
    # workload (profiles/r2b_step_metrics_summary.txt: GEMMs 23.04 GB over 205 launches; attention fwd + dQ + dK/dV + delta
    # 13.82 GB over 72 kernels = 36 calls) -- not measured by this run; only quoted for the workload it was captured on
    offline = args.workload == 'config2' and B == 64
    for key, names, label, traffic in (
            ('roofline', ['pa_gemm_tf32'], 'pa_gemm_tf32 (all projections, FFN, heads, pointer scores: fwd + dX + dW)',
             23.04e9 / 205 if offline else None),
            ('roofline_attention', ['pa_attn_fwd', 'pa_attn_bwd'], 'pa_attn_fwd + pa_attn_bwd (delta + dQ + dK/dV)',
             13.82e9 / 36 if offline else None)):
        n_l, t_ms, fl = family(names)
        if n_l:
            ach = fl / (t_ms / 1e3) / 1e12
            res[key] = {'bound': 'tensor', 'kernel': label, 'achieved': ach, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s',
                        'frac': ach / pk['tf_sustained'], 'peak_tf32': pk['tf_sustained'] / 2, 'frac_tf32': ach / (pk['tf_sustained'] / 2),
                        'traffic': traffic, 'traffic_source': 'offline: ncu pass over one step, profiles/r2b_step_metrics_summary.txt (bytes per launch)' if traffic else None,
                        'peak_source': pk['src'] + ': sustained bf16 cuBLAS; a kind::tf32 kernel tops out at half of it',
                        'avg_launch_ms': t_ms / n_l, 'launches_per_step': n_l / prof_steps, 'ms_per_step': t_ms / prof_steps,
                        'flop_per_launch': fl / n_l, 'share_of_step': (t_ms / prof_steps) / ms_step,
                        'how': f'CUDA events around every call for {prof_steps} extra steps after the headline timing'}
    if not args.no_decode:
        res['decode'] = run_decode(model, cfg, dev, world, rank, timed, args, args.workload)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res['cpu_baseline'] = cpu_baseline_train(args.workload, sample_batch=16, steps=2)     # ~15-25 s of host work
    if rank == 0 and world == 1 and not args.no_torch_cuda:
        del model, opt, step
        torch.cuda.empty_cache()
        res['torch_cuda'] = torch_cuda_arm(args.workload, B, dev, decode=not args.no_decode)
    return res


def run_decode(model, cfg, dev, world, rank, timed, args, workload):
    """BASELINE configs[2]: KV-cached greedy decode of `--decode-drawings` synthetic drawings per GPU in batches of
    `--decode-batch`, max_len = MAX_OUTPUT_LENGTH.  tokens = generated positions (every row decodes until all rows of its
    batch have emitted END, as the reference does; with the seeded-init weights no row ever emits END: 256 steps)."""
    from plankassembly_b200 import synthetic as syn
    B = args.decode_batch
    n_batches = max(1, (args.decode_drawings + B - 1) // B)
    n_distinct = min(n_batches, 4)
    host = [syn.batch_for(cfg, range((rank * 64 + i) * B + 100000, (rank * 64 + i + 1) * B + 100000)) for i in range(n_distinct)]
    host = [{k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in hb.items()} for hb in host]
    resident = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in hb.items()} for hb in host]
    # decode with the seeded-init weights (the timed train steps above have moved `model`'s weights; a
    # half-trained model may emit END everywhere at once, which would stop the loop after one step)
    model.load_state_dict(syn.init_state_dict(cfg))
    model.eval()
    from plankassembly_b200 import _lib
    with torch.no_grad():
        for _ in range(2):                             # warm-up: buffers, kernel attributes, CUDA-graph capture, allocator pools
            out = model(resident[0])
        n_tok = [0]

        def dec(i):
            o = model(resident[i % n_distinct])
            n_tok[0] += o['samples'].numel()

        l0 = _lib.launch_count()
        ms = timed(dec, n_batches)
        launches = _lib.launch_count() - l0
        tok_resident = n_tok[0] * world
        n_tok[0] = 0

        def dec_e2e(i):
            hb = host[i % n_distinct]
            batch = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in hb.items()}
            o = model(batch)
            n_tok[0] += o['samples'].cpu().numel() + o['attach'].cpu().numel() * 0

        ms_e2e = timed(dec_e2e, n_batches)
    model.train()
    Bo, T = out['samples'].shape
    d, S, L = cfg.MODEL.NUM_MODEL, cfg.DATA.MAX_INPUT_LENGTH - 1, cfg.MODEL.NUM_DECODER_LAYERS
    # algorithmic HBM bytes per generated token per sequence: cross K/V + average self K/V, fp32
    bytes_tok = 2 * L * d * (S + T / 2) * 4
    gbs = bytes_tok * tok_resident / world / (ms / 1e3) / 1e9
    pk = peaks()
    res = {'metric': 'greedy-decode generated tokens/sec (KV cache)', 'value': tok_resident / (ms / 1e3),
           'unit': UNIT, 'ms_per_decode': ms / n_batches, 'ms_per_token_step': ms / n_batches / T, 'batch_per_gpu': Bo,
           'drawings_per_gpu': n_batches * Bo, 'steps_per_decode': T,
           'engine': getattr(model._decoder_engine, 'mode', None), 'dtype': 'f32 (3xTF32 error-compensated tensor-core GEMMs where stated in DESIGN.md)',
           'gpu_launches': launches,
           'config': {'workload': f'BASELINE configs[2]: greedy decode of {n_batches * Bo} synthetic drawings per GPU in batches of {Bo}, '
                                  f'max_len {T}, S={S}, full model, seeded-init weights (no early END: {T} steps per batch)'},
           'e2e': {'value': n_tok[0] * world / (ms_e2e / 1e3), 'unit': UNIT,
                   'h2d_bytes_per_step': sum(v.numel() * v.element_size() for v in host[0].values() if torch.is_tensor(v)),
                   'd2h_bytes_per_step': Bo * T * 16},
           'roofline': {'bound': 'hbm', 'kernel': 'decode_attn (cross + self K/V streams)', 'achieved': gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                        'frac': gbs / pk['hbm_gbs'], 'traffic': None, 'peak_source': pk['src'],
                        'note': 'algorithmic fp32 K/V bytes per token (cross + mean self cache) x tokens / whole decode time per GPU '
                                '(prefill and every non-attention kernel of the step included in the time)'}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res['cpu_baseline'] = cpu_baseline_decode(workload, sample_batch=4)
    return res


# ------------------------------------------------------------------------------------------------ CPU arms
def _cpu_reference_model(workload, dropout):
    """-> (model with train_step/eval via forward, kind).  The unmodified reference when staged, else the oracle port."""
    from plankassembly_b200 import synthetic as syn
    cfg = workload_cfg(workload, dropout)
    ref = reference_models()
    if ref is not None:
        torch.manual_seed(2022)
        m = ref.build_model(cfg)
        m.load_state_dict(syn.init_state_dict(cfg))
        return m, cfg, 'reference'
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    from plank_oracle import OraclePlankModel
    return OraclePlankModel(cfg, syn.init_state_dict(cfg), requires_grad=True), cfg, 'port'


def cpu_train_steps(workload, sample_batch, steps, warmup):
    """Reference train step (fwd+bwd+Adam, configured dropout) on the host cores -> (tokens/s, ms per step, kind, cores)."""
    from plankassembly_b200 import synthetic as syn
    warnings.filterwarnings('ignore')
    torch.set_num_threads(os.cpu_count())
    m, cfg, kind = _cpu_reference_model(workload, 0.2)
    T = cfg.DATA.MAX_OUTPUT_LENGTH
    batch = syn.batch_for(cfg, range(sample_batch))
    if kind == 'reference':
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=cfg.LR)

        def one():
            opt.zero_grad(set_to_none=True)
            m(batch)['loss'].backward()
            opt.step()
    else:
        from plank_oracle import adam_train_step
        opt = torch.optim.Adam(m.parameters(), lr=cfg.LR)

        def one():
            adam_train_step(m, opt, batch)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return sample_batch * T * steps / dt, dt / steps * 1e3, kind, torch.get_num_threads()


def cpu_baseline_train(workload, sample_batch, steps):
    v, ms, kind, cores = cpu_train_steps(workload, sample_batch, steps, warmup=1)
    return {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind, 'cpu': cpu_model(), 'ms_per_step': ms,
            'sample': f'{steps} train steps fwd+bwd+Adam at batch {sample_batch} (of 64) after 1 warm-up, same shapes, dropout 0.2, fp32 torch CPU, '
                      + ('unmodified reference PlankModel (oracle/_ref)' if kind == 'reference' else 'oracle port (oracle/_ref not staged)')}


def cpu_baseline_decode(workload, sample_batch):
    """The reference's own greedy decode (ref models.py:267-323: no KV cache, O(T^3)) on a bounded sample."""
    from plankassembly_b200 import synthetic as syn
    warnings.filterwarnings('ignore')
    torch.set_num_threads(os.cpu_count())
    m, cfg, kind = _cpu_reference_model(workload, 0.0)
    batch = syn.batch_for(cfg, range(100000, 100000 + sample_batch))
    t0 = time.perf_counter()
    with torch.no_grad():
        if kind == 'reference':
            m.eval()
            out = m(batch)
        else:
            out = m.eval_step(batch)
    dt = time.perf_counter() - t0
    n = out['samples'].numel()
    return {'value': n / dt, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': kind, 'cpu': cpu_model(),
            'sample': f'one greedy decode of {sample_batch} drawings ({out["samples"].shape[1]} steps, {n} generated tokens, {dt:.1f} s), '
                      'same shapes and seeded-init weights, fp32 torch CPU, '
                      + ('unmodified reference eval_step (oracle/_ref)' if kind == 'reference' else 'oracle port')}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation, every step a bounded sample (batch 8 of 64)."""
    sample = 8
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 2))
    v, ms, kind, cores = cpu_train_steps(args.workload, sample, steps, warmup)
    cfg = workload_cfg(args.workload)
    S, T = cfg.DATA.MAX_INPUT_LENGTH - 1, cfg.DATA.MAX_OUTPUT_LENGTH
    cb = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': kind, 'cpu': cpu_model(),
          'sample': f'{steps} timed train steps fwd+bwd+Adam, each a bounded sample of the workload: batch {sample} of {args.batch} '
                    f'(tokens/s = {sample}*{T}/step time), {warmup} warm-up, dropout 0.2, fp32 torch CPU, '
                    + ('unmodified reference PlankModel (oracle/_ref)' if kind == 'reference' else 'oracle port (oracle/_ref not staged)')}
    return {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
            'warmup': warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': workload_config(args.workload, args.batch, S, T, max(1, args.gpus)),
            'cpu_baseline': cb, 'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}


# ------------------------------------------------------------------------------------------------ torch-CUDA arm
def torch_cuda_arm(workload, B, dev, decode=True, steps=5, warmup=2):
    """The unmodified reference model on the SAME B200 through torch's own CUDA kernels (cuBLAS, SDPA, native LayerNorm):
    the bar a user gets by just moving the reference to the GPU.  Train step with TF32 matmuls allowed and not; the
    reference's greedy decode (no KV cache) on one batch."""
    from plankassembly_b200 import synthetic as syn
    ref = reference_models()
    if ref is None:
        return {'unavailable': 'oracle/_ref not staged'}
    warnings.filterwarnings('ignore')
    cfg = workload_cfg(workload)
    T = cfg.DATA.MAX_OUTPUT_LENGTH
    out = {}
    batches = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in syn.batch_for(cfg, range(i * B, (i + 1) * B)).items()} for i in range(2)]
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        for label, tf in (('tf32', True), ('fp32', False)):
            torch.backends.cuda.matmul.allow_tf32 = tf
            torch.backends.cudnn.allow_tf32 = tf
            torch.manual_seed(2022)
            m = ref.build_model(cfg)
            m.load_state_dict(syn.init_state_dict(cfg))
            m = m.to(dev).train()
            opt = torch.optim.Adam(m.parameters(), lr=cfg.LR, fused=True)

            def one(i):
                opt.zero_grad(set_to_none=True)
                m(batches[i % 2])['loss'].backward()
                opt.step()
            for i in range(warmup):
                one(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                one(i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out['train_' + label] = {'value': B * T / (ms / 1e3), 'unit': UNIT, 'ms_per_step': ms, 'steps': steps,
                                     'peak_mem_gb': torch.cuda.max_memory_allocated(dev) / 2**30}
            if decode and not tf:
                m.eval()
                db = {k: (v[:64] if torch.is_tensor(v) else v[:64]) for k, v in batches[0].items()}
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                with torch.no_grad():
                    o = m(db)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                out['decode_fp32'] = {'value': o['samples'].numel() / dt, 'unit': UNIT, 'seconds': dt, 'batch': 64, 'steps': o['samples'].shape[1],
                                      'note': 'reference eval_step as is (no KV cache, host syncs per step), one batch of 64, wall clock'}
            del m, opt
            torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    out['what'] = 'unmodified reference PlankModel (oracle/_ref) on this GPU via torch ' + torch.__version__ + ' CUDA kernels, same batch/shapes/dropout, fused Adam'
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference', 'torch_cuda'])
    ap.add_argument('--workload', default='config2', choices=sorted(WORKLOADS))
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-torch-cuda', action='store_true')
    ap.add_argument('--no-decode', action='store_true')
    ap.add_argument('--torch-adam', action='store_true', help='step with torch.optim.Adam(fused=True) instead of optim.FusedAdam')
    ap.add_argument('--decode-drawings', type=int, default=1000, help='drawings per GPU in the greedy-decode leg (BASELINE configs[2]: 1000)')
    ap.add_argument('--decode-batch', type=int, default=1024, help='sequences decoded together per GPU (chains of 128 on parallel graph branches)')
    ap.add_argument('--profile-decode', action='store_true', help='bracket one greedy decode with cudaProfilerStart/Stop and exit')
    ap.add_argument('--profile-step', action='store_true', help='bracket ONE step with cudaProfilerStart/Stop (for ncu --profile-from-start off) and exit')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    rank, world, local_rank = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))

    if args.impl == 'reference':
        if rank == 0:
            print(json.dumps(run_reference(args)), flush=True)
        return
    if args.impl == 'torch_cuda':
        if rank == 0:
            dev = torch.device('cuda', local_rank)
            print(json.dumps({'impl': 'torch_cuda', 'metric': METRIC, 'unit': UNIT, **torch_cuda_arm(args.workload, args.batch, dev)}), flush=True)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    res = run_train(args, rank, world, local_rank)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
