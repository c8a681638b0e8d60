"""Build libaxisafe_b200.so in-tree with nvcc for sm_100a (no JIT cache, no setup.py)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ['api.cu', 'assemble.cu', 'eig_small.cu', 'eig_reduce.cu', 'eig_backx.cu', 'eig_hqrwin.cu', 'eig_big.cu', 'track.cu']
HEADERS = ['common.cuh', os.path.join('..', '..', 'include', 'axisafe_b200.h')]
OUT = os.path.join(HERE, 'axisafe_b200', '_lib', 'libaxisafe_b200.so')
FLAGS = (['-DOC_TIMING'] if os.environ.get('AXS_OC_TIMING') else []) + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
         '-Xcompiler', '-fPIC', '-shared']


def stale():
    if not os.path.exists(OUT):
        return True
    built = os.path.getmtime(OUT)
    deps = [os.path.join(HERE, 'csrc', s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > built for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = [nvcc] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + \
        [os.path.join(HERE, 'csrc', s) for s in SOURCES] + ['-o', OUT]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed building libaxisafe_b200.so')
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
