"""graphical-edmd_b200 -- B200 (sm_100a) implementation of Graphical-EDMD's
data-parallel hot path (whole-system prediction sweep, g(r), psi6).

The product is the C-ABI shared library ``libedmd_cuda.so`` (sources in
``csrc/``, interface in ``include/edmd_cuda.h``) and the C host in ``host/``.
This Python package is only a ctypes mirror of that ABI (``binding``) plus the
synthetic-input generator (``synth``) used by tests and ``bench.py``.

The directory name contains a hyphen, so it is loaded by path -- see
``__graft_entry__.load_package()``.
"""
from . import binding, slab, synth  # noqa: F401
from .binding import EdmdCuda, EdmdError, EdmdMg, load_library  # noqa: F401
