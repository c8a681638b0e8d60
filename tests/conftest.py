import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'axisym-safe-python_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
GOLDEN_CASES = [('pipe_soil_pml', 0, {}), ('free_steel_pipe', 0, {'nf': 60}), ('free_steel_pipe', 1, {'nf': 60}),
                ('aristegui', 0, {'nf': 40}), ('muggleton_submerged', 0, {'nf': 40}),
                ('muggleton_buried', 0, {'nf': 25}), ('gravenkamp', 0, {'nf': 30}), ('gravenkamp', 1, {'nf': 30}),
                ('gravenkamp', 2, {'nf': 30}), ('nguyen', 0, {'nf': 30}), ('nguyen', 1, {'nf': 30}),
                ('pipe_soil_pml', 1, {'nf': 20})]
CASE_IDS = ['%s-n%d' % (c[0], c[1]) for c in GOLDEN_CASES]


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_golden(name, n):
    return np.load(os.path.join(GOLDEN, '%s_n%d.npz' % (name, n)), allow_pickle=False)


def golden_dense(g, key):
    N = int(g['N'])
    out = np.zeros([N, N], complex)
    if 'G_%s_r' % key in g.files:
        np.add.at(out, (g['G_%s_r' % key], g['G_%s_c' % key]), g['G_%s_d' % key])
    return out


def make_case(name, n, kw):
    from axisafe_b200 import cases
    return cases.ALL[name](n=n, **kw)


@pytest.fixture(scope='session')
def built_lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location('axisafe_b200_build',
                                                  os.path.join(ROOT, 'axisym-safe-python_b200', 'build.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()
