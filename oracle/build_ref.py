"""TEST / BENCHMARK INFRASTRUCTURE ONLY - puts the UNMODIFIED reference package next to a matplotlib
stub under ``oracle/_ref/`` so that it can travel to the GPU box (``oracle/_ref/`` is git-ignored but
not gpurun-ignored, like the built .so files).

    python oracle/build_ref.py            # needs /root/reference (the build container)

The reference (michalkalkowski/axisym-safe-python) is 2.7 kLoC of pure Python without a setup.py, so
"building" it is a file copy of ``axisafe/*.py`` (nothing is edited) plus two empty stub modules for
``matplotlib`` / ``matplotlib.pyplot``, which the reference imports at module top (elements.py:18,
mesh.py:22, solver.py:22) and which is not installed in this image.  Used by ``bench.py --impl
reference`` and ``bench.py``'s ``cpu_baseline`` leg (``kind: "reference"``): they time the reference's own
``WaveElementAxisym.solve``.  Nothing under ``oracle/_ref`` is imported by the product.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('AXISAFE_REFERENCE_ROOT', '/root/reference')
OUT = os.path.join(HERE, '_ref')


def available():
    return os.path.isfile(os.path.join(OUT, 'axisafe', 'solver.py'))


def build():
    src = os.path.join(REF, 'axisafe')
    if not os.path.isdir(src):
        return available()
    dst = os.path.join(OUT, 'axisafe')
    os.makedirs(dst, exist_ok=True)
    for name in sorted(os.listdir(src)):
        if name.endswith('.py'):
            shutil.copyfile(os.path.join(src, name), os.path.join(dst, name))
    stub = os.path.join(OUT, 'matplotlib')
    os.makedirs(stub, exist_ok=True)
    for name in ('__init__.py', 'pyplot.py'):
        with open(os.path.join(stub, name), 'w') as fh:
            fh.write('"""empty stand-in: the reference imports matplotlib at module top, plotting is out of scope"""\n')
    return True


def load():
    """Import the reference package from oracle/_ref (raises ImportError when it was not built)."""
    if not available():
        raise ImportError('oracle/_ref not built (python oracle/build_ref.py inside the build container)')
    if OUT not in sys.path:
        sys.path.insert(0, OUT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import axisafe                      # noqa: F401
        import axisafe.mesh                 # noqa: F401
        import axisafe.solver               # noqa: F401
    return sys.modules['axisafe']


if __name__ == '__main__':
    print('oracle/_ref built' if build() else 'reference tree not found at ' + REF)
