"""TEST INFRASTRUCTURE ONLY - imports the unmodified reference from /root/reference.

Works only inside the build container (the reference tree does not travel to the
GPU box).  matplotlib is not installed and the reference imports it at module top
(elements.py:18, mesh.py:22, solver.py:22), so empty stub modules are registered
first.  Used by ``oracle/make_golden.py`` to produce the committed fixtures under
``tests/golden/`` and by container-only tests that pin ``oracle/safe_oracle.py``.
"""
import contextlib
import io
import os
import sys
import types
import warnings

REFERENCE_ROOT = os.environ.get('AXISAFE_REFERENCE_ROOT', '/root/reference')


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'axisafe'))


def load():
    """Return the reference ``axisafe`` package (imports it on first call)."""
    if 'axisafe' in sys.modules:
        return sys.modules['axisafe']
    if not available():
        raise ImportError('reference tree not present at ' + REFERENCE_ROOT)
    for name in ('matplotlib', 'matplotlib.pyplot'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import axisafe  # noqa: F401
        import axisafe.elements  # noqa: F401
        import axisafe.shape_functions  # noqa: F401
    return sys.modules['axisafe']


@contextlib.contextmanager
def quiet():
    """Silence the reference's per-frequency prints and NumPy deprecation noise."""
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter('ignore')
        yield
