"""axisafe_b200 - B200-native drop-in for the hot path of ``axisafe``.

Same module names and call signatures as the reference package
(``mesh``, ``misc``, ``solver``, plus ``elements`` and ``shape_functions``);
the per-frequency SAFE dispersion sweep runs in hand-written sm_100a CUDA
kernels behind a C-ABI (``include/axisafe_b200.h``).  See DESIGN.md.
"""
from . import mesh
from . import misc
from . import solver

__all__ = ['mesh', 'elements', 'shape_functions', 'solver', 'misc', 'cases']
