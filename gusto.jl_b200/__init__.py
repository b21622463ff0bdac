"""gusto-b200: B200-native batched GuSTO SCP hot path behind the GuSTO.jl plugin surface.

The directory name (`gusto.jl_b200`) is not a valid Python identifier; load it with
`__graft_entry__.load_package()` which registers it as the module `gusto_b200`.
"""
from . import models, problems, trajio  # noqa: F401


def engine():
    """Lazy import of the ctypes binding over the C-ABI shared library (fails loudly if it is not built)."""
    from . import host
    return host
