"""ctypes binding of the C-ABI library (include/iskra_b200.h).  No CPU fallback: if the CUDA
library is missing or no GPU is present, calls fail loudly."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libiskra_b200.so")

OK = 0
E_INVALID, E_CUDA, E_CAPACITY, E_PMAX, E_PK, E_OOB, E_NCCL, E_UNSUPPORTED, E_SINGULAR = range(-1, -10, -1)
BND_NONE, BND_WRAP, BND_DISCARD = 0, 1, 2
BC_OPEN, BC_PERIODIC = 0, 1
PUSHER_XY, PUSHER_RZ = 0, 1
EDGE_LEFT, EDGE_RIGHT, EDGE_BOTTOM, EDGE_TOP = 0, 1, 2, 3
SURF_PERIODIC, SURF_ABSORBING, SURF_REFLECTIVE, SURF_ELECTRODE_FIXED, SURF_ELECTRODE_FLOATING = range(5)
MCC_ELASTIC_ISOTROPIC, MCC_ELASTIC_BACKWARD, MCC_INELASTIC_BACKWARD, MCC_EXCITATION, MCC_IONIZATION = range(5)


class IskraError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("iskra_b200 error %d: %s" % (code, msg))
        self.code = code


vp = C.c_void_p
i32, i64, f64, u64 = C.c_int32, C.c_int64, C.c_double, C.c_uint64
dp = C.POINTER(C.c_double)

# name -> argtypes  (every function returns int32 unless listed in _RESTYPES)
SIGNATURES = {
    "iskb_version": [],
    "iskb_last_error": [],
    "iskb_create": [i32, C.POINTER(vp)],
    "iskb_destroy": [vp],
    "iskb_set_stream": [vp, vp],
    "iskb_synchronize": [vp],
    "iskb_launch_count": [vp, C.POINTER(i64)],
    "iskb_profile_enable": [vp, i32],
    "iskb_profile_read": [vp, C.POINTER(f64), C.POINTER(i64)],
    "iskb_comm_unique_id": [vp],
    "iskb_comm_init": [vp, i32, i32, vp],
    "iskb_grid_set": [vp, i32, i32, f64, f64, f64, f64, C.POINTER(i32)],
    "iskb_cell_volume": [vp, vp],
    "iskb_poisson_create": [vp, f64],
    "iskb_poisson_apply_periodic": [vp, i32],
    "iskb_poisson_apply_dirichlet": [vp, vp, f64],
    "iskb_poisson_apply_dirichlet_edge": [vp, i32, f64],
    "iskb_poisson_get_dense": [vp, vp, vp],
    "iskb_poisson_mode": [vp, C.POINTER(i32)],
    "iskb_field_solve": [vp],
    "iskb_fields_download": [vp, vp, vp, vp],
    "iskb_fields_upload": [vp, vp, vp, vp],
    "iskb_species_create": [vp, i64, f64, f64, f64, C.POINTER(vp)],
    "iskb_species_upload": [vp, vp, vp, vp, vp, i64, i64],
    "iskb_species_download": [vp, vp, vp, vp, vp, i64],
    "iskb_species_np": [vp, C.POINTER(i64)],
    "iskb_species_window_stats": [vp, vp],
    "iskb_species_sample_maxwellian": [vp, i64, vp, vp, vp, vp, u64],
    "iskb_species_copy_positions": [vp, vp, vp],
    "iskb_species_density_download": [vp, vp],
    "iskb_species_remove": [vp, i64],
    "iskb_species_add": [vp, vp],
    "iskb_species_remove_in_cells": [vp, vp, C.POINTER(i64)],
    "iskb_cell_index": [vp, vp, vp, vp, vp],
    "iskb_sort_by_cell": [vp, vp],
    "iskb_sort_for_deposit": [vp, vp],
    "iskb_gather": [vp, vp],
    "iskb_push": [vp, vp, f64],
    "iskb_boundary": [vp, i32, i32, C.POINTER(i64)],
    "iskb_density": [vp, vp],
    "iskb_rho_zero": [vp],
    "iskb_rho_accumulate": [vp, vp],
    "iskb_rho_allreduce": [vp],
    "iskb_set_after_push": [vp, i32, i32],
    "iskb_set_sort_interval": [vp, i32],
    "iskb_set_sort_policy": [vp, f64, i32],
    "iskb_set_sort_full_interval": [vp, i32],
    "iskb_set_advance_path": [vp, i32],
    "iskb_see_emit": [vp, vp, i32, vp, u64, vp],
    "iskb_set_lean": [vp, i32],
    "iskb_step_set_active": [vp, vp, i32, vp, i32],
    "iskb_ctx_counts": [vp, vp, vp, vp],
    "iskb_species_sort_stats": [vp, vp],
    "iskb_step": [vp, f64, i32],
    "iskb_stream_join": [vp],
    "iskb_mcc_create": [vp, vp, f64, f64, f64, vp, i32, vp, vp, vp, vp, vp, vp, u64, C.POINTER(vp)],
    "iskb_mcc_constants": [vp, C.POINTER(f64), C.POINTER(f64)],
    "iskb_mcc_perform": [vp, f64, vp, C.POINTER(i64), C.POINTER(i64)],
    "iskb_mcc_totals": [vp, vp],
    "iskb_poisson_add_dof": [vp, C.POINTER(i32)],
    "iskb_poisson_apply_neumann": [vp, vp, i32],
    "iskb_poisson_sigma_set": [vp, i32, f64],
    "iskb_poisson_sigma_add": [vp, i32, f64],
    "iskb_poisson_sigma_get": [vp, i32, C.POINTER(f64)],
    "iskb_poisson_dense_size": [vp, C.POINTER(i64)],
    "iskb_phi_at": [vp, i32, i32, C.POINTER(f64)],
    "iskb_tracker_create": [vp, i32, C.POINTER(vp)],
    "iskb_tracker_track_surface": [vp, vp, i32, i32, f64, C.POINTER(i32)],
    "iskb_tracker_lookup": [vp, i32, i32, i32, i32, C.POINTER(i32)],
    "iskb_tracker_track": [vp, vp, f64, C.POINTER(i64)],
    "iskb_tracker_check": [vp, vp, f64, C.POINTER(i64), C.POINTER(i32)],
    "iskb_surface_charge": [vp, i32, C.POINTER(f64), i32],
    "iskb_tracker_route_hits_to_sigma": [vp, i32],
    "iskb_warning_too_fast": [vp, C.POINTER(i32)],
    "iskb_dsmc_create": [vp, vp, vp, vp, vp, i32, u64, C.POINTER(vp)],
    "iskb_dsmc_perform": [vp, f64, vp, C.POINTER(i64), C.POINTER(i64)],
    "iskb_cell_volume_set": [vp, vp],
    "iskb_poisson_set_dense": [vp, vp, i64],
    "iskb_set_pusher": [vp, i32],
    "iskb_transform_cylindrical": [vp, f64],
}
_RESTYPES = {"iskb_last_error": C.c_char_p}

_lib = None


def lib():
    """Loads libiskra_b200.so (building it first if the sources are newer)."""
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            from . import build as _b
            _b.build()
        L = C.CDLL(SO_PATH, mode=C.RTLD_GLOBAL)
        for name, args in SIGNATURES.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = _RESTYPES.get(name, i32)
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        raise IskraError(rc, lib().iskb_last_error().decode("utf-8", "replace"))


def ptr(a):
    """void* of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"]
    return a.ctypes.data_as(vp)


def f64a(a):
    return np.ascontiguousarray(a, dtype=np.float64)
