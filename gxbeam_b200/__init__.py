"""gxbeam_b200 — B200-native batched Newton solver behind GXBeam's residual/Jacobian/solve hot path.

Only what the path needs lives here: `csrc/` (CUDA kernels + the C-ABI of include/gxbeam_b200.h), `api.py` (ctypes
binding + the batched analysis entry points that mirror the reference's `static_analysis`), `assembly.py` (host-side
packing of the reference's data model) and `state.py` (AssemblyState post-processing).
"""
from .assembly import (Assembly, Element, PrescribedConditions, default_force_scaling, discretize_beam,  # noqa: F401
                       distributed_loads, pack_distributed_loads, pack_point_masses, pack_prescribed_conditions,
                       point_mass)
from .api import Context, GXBError, static_analysis_batch, library_path  # noqa: F401
from .state import AssemblyState, assembly_state  # noqa: F401
