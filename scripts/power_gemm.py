"""SM clock and board power while the tensor-core GEMM runs back to back for a few seconds (GPU only).
The B200 is power-managed: what a kernel reaches in a 20-launch burst is not what it sustains inside a training step."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pynvml
from plankassembly_b200 import ops
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
M, N, K = int(os.environ.get('GM', 32768)), int(os.environ.get('GN', 1536)), int(os.environ.get('GK', 512))
a = torch.randn(M, K, device='cuda'); b = torch.randn(N, K, device='cuda'); c = torch.zeros(M, N, device='cuda'); bias = torch.randn(N, device='cuda')
SECS = float(os.environ.get('SECS', 2.0))


def sample(stop, out):
    while not stop.is_set():
        out.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
        time.sleep(0.02)


def run(tag, fn, flop, **env):
    for k, v in env.items():
        os.environ[k] = str(v)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    stop, out = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, out)); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    t0 = time.time(); e0.record()
    while time.time() - t0 < SECS:
        for _ in range(200): fn()
        n += 200
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    us = e0.elapsed_time(e1) / n * 1e3
    half = out[len(out) // 2:]
    clk = sorted(x[0] for x in half)[len(half) // 2]; pw = sorted(x[1] for x in half)[len(half) // 2]
    print(f'{tag:44s} {us:7.1f} us {flop / us / 1e6:7.1f} TFLOP/s | second half: SM {clk} MHz, {pw:.0f} W (max {max(x[1] for x in out):.0f} W)', flush=True)
    for k in env:
        os.environ.pop(k, None)
    time.sleep(1.0)


g = lambda: ops.gemm_tf32(a, b, c, M, N, K, lda=K, ldb=K, ldc=N, bias=bias)
print(f'M={M} N={N} K={K}, {SECS} s each')
run('ours EPI=0', g, 2 * M * N * K, PLANK_B200_GEMM_EPI=0)
run('ours EPI=2', g, 2 * M * N * K, PLANK_B200_GEMM_EPI=2)
run('ours mainloop only', g, 2 * M * N * K, PLANK_B200_GEMM_EPI=0, PLANK_B200_GEMM_DEBUG=1)
run('ours EPI=0 single-CTA', g, 2 * M * N * K, PLANK_B200_GEMM_EPI=0, PLANK_B200_GEMM_PAIR=0)
torch.backends.cuda.matmul.allow_tf32 = True
run('cuBLAS tf32 addmm', lambda: torch.addmm(bias, a, b.t()), 2 * M * N * K)
ah, bh = a.bfloat16(), b.bfloat16()
run('cuBLAS bf16 matmul', lambda: torch.matmul(ah, bh.t()), 2 * M * N * K)
x = torch.randn(64 * 1024 * 1024, device='cuda'); y = torch.empty_like(x)
run('copy 256 MB (bytes as "flop")', lambda: y.copy_(x), 2 * x.numel() * 4)
