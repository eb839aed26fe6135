"""Build the CUDA engine in-tree: localuf_b200/csrc/libluf_b200.so (sm_100a)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(CSRC, 'libluf_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
         '-shared', '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--use_fast_math']


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h', '.inl'))]
    deps.append(os.path.join(os.path.dirname(HERE), 'include', 'luf_b200.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if force or stale():
        cmd = [NVCC, *FLAGS, '-o', LIB, *sources()]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))
