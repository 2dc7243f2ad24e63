"""cgenie_b200 -- B200-native ensemble engine for cGENIE's tracer hot path.

The product is the CUDA library `libcgenie_b200.so` behind the C-ABI of
`include/cgenie_b200.h`; this package is the thin Python host that mirrors the
reference's module entry points (same names, argument meaning and error
behaviour) for tests, benchmarks and Python users.  There is no CPU fallback:
importing works anywhere, but every compute call fails loudly without the
compiled library and a CUDA device.
"""
from .engine import CgenieError, Ensemble, EnsembleGroups, TracerStep  # noqa: F401
from .jobdir import materialise  # noqa: F401

__all__ = ["Ensemble", "EnsembleGroups", "TracerStep", "CgenieError", "materialise"]
