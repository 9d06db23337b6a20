"""Build the C-ABI CUDA library in-tree:  python -m plankassembly_b200.build

nvcc cross-compiles for sm_100a without a GPU; the resulting
plankassembly_b200/csrc/libplank_b200.so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(CSRC, 'build')
LIB = os.path.join(CSRC, 'libplank_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-Xcompiler', '-fPIC', '--expt-relaxed-constexpr', '-Xptxas', '-v']
FLAGS += os.environ.get('PLANK_B200_NVCC_FLAGS', '').split()      # e.g. -DPA_ATTN_TRACE for the debug timeline


def _digest(path, deps):
    h = hashlib.sha256(' '.join(FLAGS).encode())
    for p in [path] + deps:
        with open(p, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def _compile(src, deps, verbose):
    obj = os.path.join(OBJ, os.path.basename(src) + '.o')
    stamp = obj + '.sha'
    dig = _digest(src, deps)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj, ''
    r = subprocess.run([NVCC, *FLAGS, '-c', src, '-o', obj], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f'nvcc failed for {src}:\n{r.stdout}\n{r.stderr}')
    with open(stamp, 'w') as f:
        f.write(dig)
    return obj, r.stderr


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cuh'))
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'plank_b200.h'))
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(s, deps, verbose), srcs))
    objs = [o for o, _ in results]
    log = '\n'.join(l for _, l in results if l)
    if verbose and log:
        print(log)
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        r = subprocess.run([NVCC, '-shared', '-o', LIB, *objs, '-gencode', 'arch=compute_100a,code=sm_100a',
                            '-cudart', 'static'], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f'link failed:\n{r.stdout}\n{r.stderr}')
    return LIB


if __name__ == '__main__':
    path = build(verbose='-v' in sys.argv, force='-f' in sys.argv)
    print(path)
