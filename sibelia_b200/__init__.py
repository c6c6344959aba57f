"""sibelia_b200 -- B200-native de Bruijn-graph hot path of Sibelia (bifurcation enumeration + bulge removal).

The product is the C-ABI shared library ``libsibgpu.so`` (include/sibgpu.h) built from sibelia_b200/csrc/ plus the
reference-side binding (sibelia_b200/csrc/facade/: translation units that define the reference's own members on top of
the C ABI and replace vertexenumeration.cpp, blockfinder.cpp + bulgeremoval.cpp and fasta.cpp in its build).  This Python
package is only the ctypes binding used by tests/ and bench.py.  There is no CPU fallback anywhere in it.
"""
from .binding import (Context, SibgpuError, INST_DTYPE, build, lib_path, load, device_count)  # noqa: F401
