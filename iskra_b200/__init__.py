"""iskra_b200 -- B200-native particle hot path of bchaber/iskra behind the reference's module API.

Modules mirror the reference packages the problem scripts use:
    regular_grids              <- RegularGrids
    finite_difference_method   <- FiniteDifferenceMethod
    particle_in_cell           <- ParticleInCell
    chemistry                  <- Chemistry (MCC)
    circuit                    <- Circuit (series RLC, host side)
    diagnostics                <- Diagnostics (openPMD records, fetched from the device on demand)
    configuration, units_and_constants <- problem/configuration.jl, problem/units_and_constants.jl
All compute goes through the C ABI in include/iskra_b200.h (libiskra_b200.so, CUDA sm_100a).
There is no CPU fallback.
"""
from . import _lib, chemistry, circuit, configuration, diagnostics, datasets, finite_difference_method, particle_in_cell  # noqa: F401
from . import regular_grids, runtime, units_and_constants  # noqa: F401
from ._lib import IskraError  # noqa: F401
