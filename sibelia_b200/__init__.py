"""sibelia_b200 -- B200-native de Bruijn-graph hot path of Sibelia (bifurcation enumeration + bulge removal).

The product is the C-ABI shared library ``libsibgpu.so`` (include/sibgpu.h) built from sibelia_b200/csrc/ plus the
C++ facade (sibelia_b200/csrc/facade/, namespace SyntenyFinder) that mirrors the reference's classes.  This Python
package is only the ctypes binding used by tests/ and bench.py.  There is no CPU fallback anywhere in it.
"""
from .binding import (Context, SibgpuError, INST_DTYPE, build, lib_path, load, device_count)  # noqa: F401
