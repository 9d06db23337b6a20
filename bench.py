"""bench.py -- shape-program tokens/sec of the PlankAssembly hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--decode-drawings D] [--no-decode]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one synthetic batch:
  train  (default, BASELINE.json configs[1]): full model d=512 6+6 layers, per-GPU batch 64,
         S=512 encoder / T=256 decoder positions, configured dropout 0.2, forward + backward +
         gradient all-reduce (N>1) + Adam.  tokens = decoder positions B*T (the "shape program").
  decode (configs[2], nested under "decode"): KV-cached greedy decode of D drawings per GPU in batches of 64, max_len 256
         (default D = 128; BASELINE names 1000); tokens = generated positions.
Prints ONE JSON line (rank 0).  `value` = device-resident inputs; `e2e` = the same metric through
the public PlankModel API with pinned HOST batches copied in and the loss read back every step.
`--impl reference` times the CPU port of the reference (oracle/plank_oracle.py, see DESIGN.md for
why the reference itself cannot travel to the GPU box) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'shape-program tokens/sec (train step, fwd+bwd+Adam)'
UNIT = 'tokens/s'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tf_burst': d['bf16_tflops'], 'tf_sustained': d['bf16_tflops_sustained'], 'src': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'src': 'fallback'}


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
def attention_flops_fwd(B, H, dh, S, T, L_enc, L_dec):
    """Algorithmic attention FLOPs of one forward (full, not causal-halved)."""
    d = H * dh
    return B * (L_enc * S * 4 * S * d + L_dec * T * 4 * T * d + L_dec * T * 4 * S * d)


def run_train(args, rank, world, local_rank):
    import torch.distributed as dist
    from plankassembly_b200 import _lib, synthetic as syn
    from plankassembly_b200.models import build_model
    from plankassembly_b200.parallel import GradAllReduce

    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if not _lib.load().pa_device_ok():
        raise SystemExit('bench.py needs a B200 (sm_100a): ' + _lib.load().pa_last_error().decode())
    cfg = syn.config2(dropout=0.2)
    B, S, T = args.batch, cfg.DATA.MAX_INPUT_LENGTH - 1, cfg.DATA.MAX_OUTPUT_LENGTH
    torch.manual_seed(2022)
    model = build_model(cfg)
    model.load_state_dict(syn.init_state_dict(cfg))
    model = model.to(dev).train()
    # N > 1: every .grad is a view into one flat buffer -> ONE NCCL all-reduce per step (parallel.py)
    reducer = GradAllReduce(model.parameters()) if world > 1 else None
    opt = torch.optim.Adam(model.parameters(), lr=cfg.LR, fused=True)

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
    # --- device-resident timing; the dominant kernel is additionally bracketed with events
    dom_events = []
    _lib.PROFILE_HOOK = (args.dominant, dom_events)
    with ClockSampler(local_rank) as clk:
        ms = timed(lambda i: step(resident[i % n_host]), args.steps)
    _lib.PROFILE_HOOK = None
    launches = _lib.launch_count() - l0
    dom_ms = [a.elapsed_time(b) for a, b in dom_events]

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

    tokens = B * T * world
    res = {
        'metric': METRIC, 'value': tokens * args.steps / (ms / 1e3), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'BASELINE configs[1]: train_complete full model d=512 H=8 ff=1024 6+6 layers, dropout 0.2, '
                               f'per-GPU batch {B}, S={S} encoder / T={T} decoder positions, fwd+bwd+allreduce+Adam',
                   'global_batch': B * world, 'parallelism': f'dp{world}', 'attention': model.attn_impl,
                   'l2': 'per-step activations (>3 GB) exceed the 126 MB L2; no explicit flush',
                   'e2e': 'pinned H2D of every batch + D2H of (loss, accuracy) every step; the host reads step i while step i+1 runs',
                   'encoder_tokens_per_step': B * S * world},
        'clocks': clk.summary,
        'e2e': {'value': tokens * args.steps / (ms_e2e / 1e3), 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 8},
        'gpu_launches': launches,
    }
    # --- roofline of the dominant kernel (tensor bound: attention contractions)
    pk = peaks()
    if dom_ms:
        fl = attention_flops_fwd(B, cfg.MODEL.NUM_HEAD, cfg.MODEL.NUM_MODEL // cfg.MODEL.NUM_HEAD, S, T,
                                 cfg.MODEL.NUM_ENCODER_LAYERS, cfg.MODEL.NUM_DECODER_LAYERS)
        mult = {'pa_attn_fwd': 1.0, 'pa_attn_bwd': 2.5}[args.dominant]
        calls_per_step = len(dom_ms) / args.steps
        flops_per_launch = fl * mult / calls_per_step
        avg_ms = sum(dom_ms) / len(dom_ms)
        ach = flops_per_launch / (avg_ms / 1e3) / 1e12
        # DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, averaged over the 18 calls of one step) from the
        # committed ncu pass profiles/r1b_step_metrics.csv: pa_attn_bwd = delta + dQ + dK/dV kernels
        traffic = {'pa_attn_bwd': 612.8e6, 'pa_attn_fwd': 187.2e6}[args.dominant] if B == 64 else None
        res['roofline'] = {'bound': 'tensor', 'kernel': args.dominant, 'achieved': ach, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s',
                           'frac': ach / pk['tf_sustained'], 'traffic': traffic, 'traffic_unit': 'bytes per launch (ncu, profiles/r1b_step_metrics_summary.txt)',
                           'peak_source': pk['src'] + ' (sustained bf16)',
                           'avg_launch_ms': avg_ms, 'launches_per_step': calls_per_step,
                           'share_of_step': sum(dom_ms) / ms}
    if not args.no_decode:
        res['decode'] = run_decode(model, cfg, host, resident, dev, world, rank, timed, args.decode_drawings)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res['cpu_baseline'] = cpu_baseline_train(sample_batch=4, steps=3)     # ~10-20 s of host work on the box's cores
    return res


def run_decode(model, cfg, host, resident, dev, world, rank, timed, n_drawings):
    """BASELINE configs[2]: KV-cached greedy decode of `n_drawings` synthetic drawings per GPU in batches of the bench batch
    size, max_len = MAX_OUTPUT_LENGTH (`--decode-drawings 1000` is the configuration BASELINE.json names).
    tokens = generated positions (every row decodes until all rows of its batch have emitted END, as the reference does)."""
    from plankassembly_b200 import synthetic as syn
    B = resident[0]['input_value'].shape[0]
    n_batches = max(1, (n_drawings + B - 1) // B)
    host, resident = list(host), list(resident)
    while len(host) < n_batches:                           # more drawings than the train loop's four batches
        i = len(host)
        hb = syn.batch_for(cfg, range((rank * 64 + i) * B + 100000, (rank * 64 + i + 1) * B + 100000))
        hb = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in hb.items()}
        host.append(hb)
        resident.append({k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in hb.items()})
    # decode with the seeded-init weights (the timed train steps above have moved `model`'s weights; a
    # half-trained model may emit END everywhere at once, which would stop the loop after one step)
    model.load_state_dict(syn.init_state_dict(cfg))
    model.eval()
    with torch.no_grad():
        out = model(resident[0])                       # warm-up: buffers, cuBLAS workspaces, CUDA-graph capture
        n_tok = [0]

        def dec(i):
            o = model(resident[i % n_batches])
            n_tok[0] += o['samples'].numel()

        ms = timed(dec, n_batches)
        tok_resident = n_tok[0] * world
        n_tok[0] = 0

        def dec_e2e(i):
            hb = host[i % n_batches]
            batch = {k: (v.to(dev, non_blocking=True) if torch.is_tensor(v) else v) for k, v in hb.items()}
            o = model(batch)
            n_tok[0] += o['samples'].cpu().numel() + o['attach'].cpu().numel() * 0

        ms_e2e = timed(dec_e2e, n_batches)
    model.train()
    B, T = out['samples'].shape
    d, S, L = cfg.MODEL.NUM_MODEL, cfg.DATA.MAX_INPUT_LENGTH - 1, cfg.MODEL.NUM_DECODER_LAYERS
    # algorithmic HBM bytes per generated token per sequence: cross K/V + average self K/V, fp32
    bytes_tok = 2 * L * d * (S + T / 2) * 4
    gbs = bytes_tok * tok_resident / world / (ms / 1e3) / 1e9
    pk = peaks()
    return {'metric': 'greedy-decode generated tokens/sec (KV cache, CUDA-graph step)', 'value': tok_resident / (ms / 1e3),
            'unit': UNIT, 'ms_per_decode': ms / n_batches, 'batch_per_gpu': B, 'drawings_per_gpu': n_batches * B, 'steps_per_decode': T,
            'engine': getattr(model._decoder_engine, 'mode', None),
            'e2e': {'value': n_tok[0] * world / (ms_e2e / 1e3), 'unit': UNIT},
            'roofline': {'bound': 'hbm', 'achieved': gbs, 'peak': pk['hbm_gbs'], 'unit': 'GB/s', 'frac': gbs / pk['hbm_gbs'],
                         'note': 'algorithmic fp32 K/V bytes per token (cross + mean self cache) / measured time per GPU'}}


def cpu_baseline_train(sample_batch, steps):
    """CPU port of the reference (oracle) on the host cores: bounded sample of the same workload."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    from plank_oracle import OraclePlankModel, adam_train_step
    from plankassembly_b200 import synthetic as syn
    torch.set_num_threads(os.cpu_count())
    cfg = syn.config2(dropout=0.2)
    T = cfg.DATA.MAX_OUTPUT_LENGTH
    m = OraclePlankModel(cfg, syn.init_state_dict(cfg), requires_grad=True)
    opt = torch.optim.Adam(m.parameters(), lr=cfg.LR)
    batch = syn.batch_for(cfg, range(sample_batch))
    adam_train_step(m, opt, batch)                       # warm-up
    t0 = time.perf_counter()
    for _ in range(steps):
        adam_train_step(m, opt, batch)
    dt = time.perf_counter() - t0
    return {'value': sample_batch * T * steps / dt, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'{steps} train step(s) fwd+bwd+Adam at batch {sample_batch} (of 64), same shapes/dropout, fp32 torch CPU'}


def run_reference(args):
    cb = cpu_baseline_train(sample_batch=4, steps=max(1, min(args.steps, 3)))
    return {'impl': 'reference', 'metric': METRIC, 'value': cb['value'], 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': {'workload': 'BASELINE configs[1] (bounded sample: batch 4 of 64)'},
            'cpu_baseline': cb, 'e2e': {'value': cb['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--dominant', default='pa_attn_bwd', choices=['pa_attn_fwd', 'pa_attn_bwd'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-decode', action='store_true')
    ap.add_argument('--decode-drawings', type=int, default=128, help='drawings per GPU in the greedy-decode leg (BASELINE configs[2]: 1000)')
    ap.add_argument('--profile-decode', action='store_true', help='bracket one greedy decode with cudaProfilerStart/Stop and exit')
    ap.add_argument('--profile-step', action='store_true', help='bracket ONE step with cudaProfilerStart/Stop (for ncu --profile-from-start off) and exit')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup
    rank, world, local_rank = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))

    if args.impl == 'reference':
        if rank == 0:
            print(json.dumps(run_reference(args)), flush=True)
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
