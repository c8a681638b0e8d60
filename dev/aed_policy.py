"""Host model of the QR cascade (tests/host/aed_host solve): per-stage macro steps and AED calls for
AED policies, on sampled cfg1 Hessenberg matrices.  Cost model from the measured per-stage kilo-cycles
(profiles/r2e_bench.json, gpurun_out/b2.json): see DESIGN.md 4.5."""
import os, subprocess, sys, struct, tempfile, re
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'axisym-safe-python_b200'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import test_host as th

def run(exe, H, nw, env):
    n = H.shape[0]
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, 'in'), os.path.join(d, 'out')
        with open(fi, 'wb') as f:
            f.write(struct.pack('5i', n, 0, n - 1, nw, 8))
            f.write(np.asarray(H, dtype=complex).tobytes(order='F'))
        e = dict(os.environ, AED_STAGES='1', **env)
        r = subprocess.run([exe, 'solve', fi, fo], check=True, capture_output=True, text=True, env=e)
        raw = open(fo, 'rb').read()
    st = struct.unpack('5i', raw[:20])
    m = re.search(r'STAGES macro (\d+) (\d+) (\d+) (\d+) aed (\d+) (\d+) (\d+) (\d+) iters (\d+)', r.stderr)
    return st, [int(x) for x in m.groups()]

if __name__ == '__main__':
    exe = th._aed_host_exe()
    idx = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else [3, 250, 700, 1000, 1400, 1900]
    pols = {'none': (0, {}), 'all nw8': (8, {}), 'all nw6': (6, {}), 'all nw7': (7, {}), 'all nw5': (5, {}),
            'nw8 nibble0': (8, {'AED_NIBBLE': '0'}), 'nw8 nibble30': (8, {'AED_NIBBLE': '30'}), 'nw8 nibble50': (8, {'AED_NIBBLE': '50'}),
            'nw8 minm16': (8, {'AED_MINM': '16'}), 'nw8 minm32': (8, {'AED_MINM': '32'}),
            'nw8 every2': (8, {'AED_EVERY': '2'}),
            'u>=117': (8, {'AED_UMIN': '117'}), 'u>=82': (8, {'AED_UMIN': '82'})}
    acc = {k: np.zeros(9) for k in pols}
    for i in idx:
        H, _ = th._safe_hessenberg(i)
        for k, (nw, env) in pols.items():
            st, stages = run(exe, H, nw, env)
            assert st[4] == 0
            acc[k] += np.array(stages, float)
    for k in pols:
        a = acc[k] / len(idx)
        step = np.array([1525., 707., 834., 501.])            # SM-cycles per macro step per stage (r2e, no AED)
        for label, call in (('shm', np.array([105e3, 53e3, 36e3, 30e3])), ('reg?', np.array([50e3, 25e3, 17e3, 13.5e3]))):
            cost = (a[:4] * step).sum() + (a[4:8] * call * (int(pols[k][0]) / 8.0) ** 2 if False else a[4:8] * call).sum()
            print('   cost %s %.0f k' % (label, cost / 1e3), end='')
        print()
        print('%-18s macro/stage %6.0f %6.0f %6.0f %6.0f | aed calls %5.1f %5.1f %5.1f %5.1f | window iters %5.0f' % ((k,) + tuple(a)))
