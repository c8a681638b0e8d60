"""Build libaxisafe_b200.so in-tree with nvcc for sm_100a (no JIT cache, no setup.py).

Every source is compiled to an object file under ``build/`` (git-ignored) in parallel and only
when it or a header changed; the objects are linked into ``axisafe_b200/_lib/libaxisafe_b200.so``.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ['api.cu', 'assemble.cu', 'eig_small.cu', 'eig_reduce.cu', 'eig_backx.cu', 'eig_big.cu', 'eig_bigred.cu', 'track.cu']
HEADERS = ['common.cuh', 'aed.cuh', 'aed_variants.cuh', 'hqr_mb.cuh', 'modes.cuh', 'reduce_oc.cuh', 'backnorm.cuh', 'assemble_kernels.cuh', 'track_kernels.cuh', 'reduce_l2.cuh', 'backx_kernels.cuh', 'bigred_kernels.cuh', 'big_kernels.cuh', 'post_kernels.cuh', os.path.join('..', '..', 'include', 'axisafe_b200.h')]
OUT = os.path.join(HERE, 'axisafe_b200', '_lib', 'libaxisafe_b200.so')
OBJ = os.path.join(HERE, 'build')
FLAGS = (['-DOC_TIMING'] if os.environ.get('AXS_OC_TIMING') else []) + (['-DAED_PROFILE'] if os.environ.get('AXS_AED_PROFILE') else []) + [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC']


def _src(name):
    return os.path.join(HERE, 'csrc', name)


def _newest_header():
    return max([os.path.getmtime(_src(h)) for h in HEADERS] + [os.path.getmtime(os.path.abspath(__file__))])


def stale():
    if not os.path.exists(OUT):
        return True
    built = os.path.getmtime(OUT)
    return any(os.path.getmtime(_src(d)) > built for d in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    hdr = _newest_header()

    def compile_one(name):
        obj = os.path.join(OBJ, name.replace('.cu', '.o'))
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(_src(name)), hdr):
            return obj, None
        cmd = [nvcc] + FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', _src(name), '-o', obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError('nvcc failed on %s\n%s%s' % (name, res.stdout, res.stderr))
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as pool:
        results = list(pool.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            if log:
                print(log)
    res = subprocess.run([nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a'] + [o for o, _ in results] +
                         ['-o', OUT], capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed linking libaxisafe_b200.so')
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
