"""Builds ``libvkocr_b200.so`` (hand-written sm_100a kernels + the C ABI) in-tree with nvcc.

Incremental: one object per ``csrc/*.cu``, rebuilt when the source or any header is newer.  The shared object
lands next to this file so that it travels to the GPU box with the repo snapshot.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OBJ = os.path.join(HERE, 'build')
LIB = os.path.join(HERE, 'libvkocr_b200.so')

NVCC_FLAGS = [
    '-std=c++17', '-O3', '-lineinfo', '-DVKOCR_PRECISE_MATH',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-Xcompiler', '-fPIC',
    '--expt-relaxed-constexpr', '-Xptxas', '-v',
]


def _nvcc() -> str:
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found; the vkocr_b200 CUDA library cannot be built')
    return nvcc


def _compile(src: str, obj: str, log_dir: str) -> str:
    cmd = [_nvcc(), *NVCC_FLAGS, '-c', src, '-o', obj]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(log_dir, os.path.basename(src) + '.log'), 'w') as f:
        f.write(' '.join(cmd) + '\n' + proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError(f'nvcc failed on {src}:\n{proc.stdout}\n{proc.stderr}')
    return obj


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith('.cu'))
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    header_mtime = max([os.path.getmtime(h) for h in headers] + [os.path.getmtime(__file__)])
    jobs = []
    objs = []
    for s in sources:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + '.o')
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), header_mtime)
        if stale:
            jobs.append((src, obj))
    if jobs:
        if verbose:
            print(f'[vkocr_b200] compiling {len(jobs)} CUDA source(s) for sm_100a', file=sys.stderr)
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            list(pool.map(lambda j: _compile(j[0], j[1], OBJ), jobs))
    if jobs or not os.path.exists(LIB):
        cmd = [_nvcc(), '-shared', '-o', LIB, *objs, '-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f'link failed:\n{proc.stdout}\n{proc.stderr}')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
