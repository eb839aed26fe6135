#!/usr/bin/env python
"""Build variants of the CUDA library (LUF_OPT_* switches of luf_core.cuh) and,
on a GPU box, time bench.py with each.

    python tools/variants.py build  NAME=-DLUF_OPT_X=1,-DLUF_OPT_Y=0 ...
    python tools/variants.py run [--workload cfg3] [--test]     # on the GPU box
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VDIR = os.path.join(ROOT, 'localuf_b200', 'csrc', 'variants')
sys.path.insert(0, ROOT)


def build(specs):
    from localuf_b200 import build as B
    os.makedirs(VDIR, exist_ok=True)
    for f in glob.glob(os.path.join(VDIR, '*.so')):
        os.remove(f)
    procs = []
    for spec in specs:
        name, _, defs = spec.partition('=')
        out = os.path.join(VDIR, f'libluf_{name}.so')
        cmd = [B.NVCC, *B.FLAGS, *[d for d in defs.split(',') if d], '-o', out, *B.sources()]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        out, _ = p.communicate()
        regs = [l for l in out.splitlines() if 'k_mc_ufILi0E' in l or 'Used' in l]
        use = ''
        for i, l in enumerate(out.splitlines()):
            if "Compiling entry function '_Z7k_mc_ufILi0E" in l:
                use = ' '.join(out.splitlines()[i + 1:i + 4])
        print(name, 'rc', p.returncode, use[-150:])


def run(argv):
    wl = argv[argv.index('--workload') + 1] if '--workload' in argv else 'cfg3'
    for lib in sorted(glob.glob(os.path.join(VDIR, '*.so'))):
        env = dict(os.environ, LUF_B200_LIB=lib)
        name = os.path.basename(lib)[7:-3]
        if '--test' in argv:
            r = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_gpu_uf.py'), '-x', '-q', '-m', 'gpu'],
                               env=env, capture_output=True, text=True, cwd=ROOT)
            print(name, 'tests:', r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-200:])
        r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--no-cpu', '--steps', '3', '--workload', wl],
                           env=env, capture_output=True, text=True, cwd=ROOT)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(f"{name:28s} {d['value'] / 1e6:8.2f} M {d['unit']}  e2e {d['e2e']['value'] / 1e6:8.2f} M", flush=True)
        except Exception:
            print(name, 'FAILED', r.stderr[-300:], flush=True)


if __name__ == '__main__':
    if sys.argv[1] == 'build':
        build(sys.argv[2:])
    else:
        run(sys.argv[2:])
