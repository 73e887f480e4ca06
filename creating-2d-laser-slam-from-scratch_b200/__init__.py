"""B200-native 2-D laser SLAM front-end hot path (Karto correlative matcher, occupancy ray-cast, GN fine-align).

The product is the C-ABI shared library built from csrc/ (include/b200slam.h); this Python package is only the
ctypes harness used by tests/ and bench.py plus the synthetic workload generator.  Import with
    importlib.import_module("creating-2d-laser-slam-from-scratch_b200")
(the directory name is not a Python identifier).
"""
from . import abi, synth  # noqa: F401



def load(name: str):
    """Import a harness submodule that binds libb200slam.so (matcher, occgrid, hector, gmapping) on demand."""
    import importlib
    return importlib.import_module(f"{__name__}.{name}")
