"""ctypes mirror of include/vx3b200.h and loader for the in-tree libvx3b200.so.

The library is the product; this module only binds it.  There is no Python or CPU
fallback: if the shared library is missing, importing the binding raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VX3B_LIB") or os.path.join(_HERE, "libvx3b200.so")  # VX3B_LIB: an experimental build (__graft_entry__.py VX3B_VARIANT)

MAX_TOKENS = 64
FP32, FP64 = 32, 64
DOF_ALL = 63


class Token(C.Structure):
    _fields_ = [("op", C.c_int32), ("pad", C.c_int32), ("value", C.c_double)]


class VoxelMaterial(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "E", "nu", "rho", "alphaCTE", "muStatic", "muKinetic", "zetaInternal", "zetaGlobal", "zetaCollision",
        "epsilonFail", "gravMult", "mass", "massInverse", "sqrtMass", "momentInertia", "momentInertiaInverse",
        "_2xSqMxExS", "_2xSqIxExSxSxS")] + [
        ("nomSize", C.c_double), ("Cilia", C.c_double)] + [(n, C.c_int32) for n in (
            "matid", "fixed", "isTarget", "isMeasured", "sticky", "linear")] + [("rgba", C.c_uint32), ("isPaceMaker", C.c_int32), ("padSignal", C.c_int32)] + [
        (n, C.c_double) for n in ("PaceMakerPeriod", "signalValueDecay", "signalTimeDelay", "inactivePeriod", "RemoveFromSimulationAfterThisManySeconds",
                                  "TurnOnThermalExpansionAfterThisManySeconds", "TurnOnCiliaAfterThisManySeconds")]


MAX_SS_POINTS = 16


class LinkMaterial(C.Structure):
    _fields_ = [(n, C.c_float) for n in (
        "E", "nu", "eHat", "a1", "a2", "b1", "b2", "b3", "sqA1", "sqA2xIp", "sqB1", "sqB2xFMp", "sqB3xIp",
        "epsilonFail")] + [("mat1", C.c_int32), ("mat2", C.c_int32), ("linear", C.c_int32), ("num_points", C.c_int32),
                           ("strain_data", C.c_float * MAX_SS_POINTS), ("stress_data", C.c_float * MAX_SS_POINTS)]


class WorldDesc(C.Structure):
    _fields_ = [
        ("num_voxels", C.c_int32), ("num_links", C.c_int32), ("num_voxel_mats", C.c_int32), ("num_link_mats", C.c_int32),
        ("vox_size", C.c_double),
        ("lattice", C.POINTER(C.c_int16)), ("voxel_mat", C.POINTER(C.c_int32)),
        ("phase_offset", C.POINTER(C.c_double)), ("base_cilia", C.POINTER(C.c_double)), ("shift_cilia", C.POINTER(C.c_double)),
        ("dof_fixed", C.POINTER(C.c_uint8)), ("ext_force", C.POINTER(C.c_float)), ("ext_moment", C.POINTER(C.c_float)),
        ("ext_translation", C.POINTER(C.c_double)), ("ext_rotation", C.POINTER(C.c_double)),
        ("link_neg", C.POINTER(C.c_int32)), ("link_pos", C.POINTER(C.c_int32)), ("link_axis", C.POINTER(C.c_uint8)),
        ("link_mat", C.POINTER(C.c_int32)),
        ("voxel_mats", C.POINTER(VoxelMaterial)), ("link_mats", C.POINTER(LinkMaterial)),
        ("dt_frac", C.c_double), ("floor_enabled", C.c_int32), ("vary_temp_enabled", C.c_int32),
        ("temp_amplitude", C.c_double), ("temp_period", C.c_double),
        ("enable_collision", C.c_int32), ("enable_cilia", C.c_int32), ("enable_expansion", C.c_int32), ("vx2_compat", C.c_int32),
        ("max_dist_in_voxel_lengths_to_count_as_pair", C.c_double),
        ("save_position_of_all_voxels", C.c_int32), ("initial_temp", C.c_float),
        ("stop_condition", Token * MAX_TOKENS), ("fitness", Token * MAX_TOKENS), ("force_field", (Token * MAX_TOKENS) * 3),
        ("ghost", C.POINTER(C.c_uint8)),
        ("record_step_size", C.c_int32), ("record_voxel", C.c_int32), ("record_link", C.c_int32),
        ("enable_signals", C.c_int32), ("enable_attach", C.c_int32), ("enable_detach", C.c_int32), ("safety_guard", C.c_int32),
        ("secondary_experiment", C.c_int32), ("pad_dyn", C.c_int32), ("reinitialize_initial_position_after", C.c_double),
        ("attach_condition", (Token * MAX_TOKENS) * 5),
    ]


class SlabInfo(C.Structure):
    _fields_ = [("part", C.c_int32), ("n_parts", C.c_int32), ("axis", C.c_int32), ("layer_lo", C.c_int32), ("layer_hi", C.c_int32), ("num_own", C.c_int32),
                ("n_send_lo", C.c_int32), ("n_send_hi", C.c_int32), ("n_ghost_lo", C.c_int32), ("n_ghost_hi", C.c_int32),
                ("body_voxels", C.c_int64), ("body_links", C.c_int64), ("body_dt", C.c_double),
                ("send_lo", C.POINTER(C.c_int32)), ("send_hi", C.POINTER(C.c_int32)), ("ghost_lo", C.POINTER(C.c_int32)), ("ghost_hi", C.POINTER(C.c_int32)),
                ("global_index", C.POINTER(C.c_int64))]


class HaloEndpoint(C.Structure):
    _fields_ = [("ipc_handle", C.c_uint8 * 64), ("device_ptr", C.c_uint64), ("device", C.c_int32), ("n", C.c_int32), ("record_bytes", C.c_int32), ("pad", C.c_int32),
                ("state_ipc_handle", (C.c_uint8 * 64) * 6), ("state_ptr", C.c_uint64 * 6)]


class WorldResult(C.Structure):
    _fields_ = [
        ("current_time", C.c_double), ("fitness_score", C.c_double), ("initial_com", C.c_double * 3), ("current_com", C.c_double * 3),
        ("total_distance_of_all_voxels", C.c_double), ("vox_size", C.c_double), ("recent_angle", C.c_double),
        ("target_closeness", C.c_double), ("dt", C.c_double), ("steps", C.c_int64),
        ("num_voxel", C.c_int32), ("num_measured_voxel", C.c_int32), ("num_close_pairs", C.c_int32), ("collision_count", C.c_int32),
        ("diverged", C.c_int32), ("stopped", C.c_int32), ("contact_rebuilds", C.c_int32), ("contact_overflows", C.c_int32),
    ]


# every symbol include/vx3b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_DP = C.POINTER(C.c_double)
SYMBOLS = {
    "vx3b_last_error": (C.c_char_p, []),
    "vx3b_version": (C.c_char_p, []),
    "vx3b_model_from_vxa": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(_P)]),
    "vx3b_model_from_files": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(_P)]),
    "vx3b_model_new": (C.c_int, [C.c_double, C.POINTER(_P)]),
    "vx3b_model_add_material": (C.c_int, [_P, C.c_double, C.c_double]),
    "vx3b_model_set_material_param": (C.c_int, [_P, C.c_int, C.c_char_p, C.c_double]),
    "vx3b_model_set_voxel": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    "vx3b_model_set_fixed": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _DP, _DP]),
    "vx3b_model_set_force": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _DP, _DP]),
    "vx3b_model_set_world_param": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "vx3b_model_finalize": (C.c_int, [_P]),
    "vx3b_model_desc": (C.c_int, [_P, C.POINTER(WorldDesc)]),
    "vx3b_model_merged_vxa": (C.c_int, [_P, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]),
    "vx3b_model_free": (None, [_P]),
    "vx3b_ctx_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(_P)]),
    "vx3b_ctx_destroy": (None, [_P]),
    "vx3b_ctx_set_option": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "vx3b_batch_upload": (C.c_int, [_P, C.POINTER(WorldDesc), C.c_int]),
    "vx3b_batch_upload_models": (C.c_int, [_P, C.POINTER(_P), C.c_int]),
    "vx3b_batch_run": (C.c_int, [_P, C.c_int64, C.c_int]),
    "vx3b_world_time_step": (C.c_int, [C.POINTER(WorldDesc), C.POINTER(C.c_double)]),
    "vx3b_batch_read_aux": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "vx3b_batch_set_force": (C.c_int, [_P, C.c_int, C.c_int, _DP, _DP]),
    "vx3b_batch_set_global_damping": (C.c_int, [_P, C.c_int, C.c_int, C.c_double]),
    "vx3b_job_history": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int]),
    "vx3b_batch_run_async": (C.c_int, [_P, C.c_int64]),
    "vx3b_batch_halo_set": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int]),
    "vx3b_batch_halo_record_bytes": (C.c_int, [_P]),
    "vx3b_batch_halo_export": (C.c_int, [_P, C.c_int, C.c_int, C.c_void_p]),
    "vx3b_batch_halo_import": (C.c_int, [_P, C.c_int, C.c_int, C.c_void_p]),
    "vx3b_batch_halo_open": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(HaloEndpoint)]),
    "vx3b_batch_halo_connect": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(HaloEndpoint), C.c_int]),
    "vx3b_batch_run_slab": (C.c_int, [_P, C.c_int64]),
    "vx3b_batch_halo_slots": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int32)]),
    "vx3b_batch_halo_connect_fused": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(HaloEndpoint), C.POINTER(C.c_int32), C.c_int]),
    "vx3b_batch_halo_timeouts": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "vx3b_batch_halo_wait_ns": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "vx3b_model_from_vxa_slab": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "vx3b_model_slab_info": (C.c_int, [_P, C.POINTER(SlabInfo)]),
    "vx3b_batch_results": (C.c_int, [_P, C.POINTER(WorldResult)]),
    "vx3b_batch_read_state": (C.c_int, [_P, C.c_int, _DP, _DP, _DP, _DP]),
    "vx3b_batch_read_initial_pos": (C.c_int, [_P, C.c_int, _DP]),
    "vx3b_batch_read_links": (C.c_int, [_P, C.c_int, _DP, _DP, _DP, C.POINTER(C.c_int32)]),
    "vx3b_batch_collision_pairs": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vx3b_batch_num_links": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32)]),
    "vx3b_batch_read_signals": (C.c_int, [_P, C.c_int, _DP, C.POINTER(C.c_int32)]),
    "vx3b_batch_read_topology": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "vx3b_ctx_timing": (C.c_int, [_P, C.POINTER(C.c_float), C.POINTER(C.c_int64)]),
    "vx3b_ctx_total_voxel_steps": (C.c_int64, [_P]),
    "vx3b_ctx_set_stream": (C.c_int, [_P, C.c_uint64]),
    "vx3b_model_set_material_bilinear": (C.c_int, [_P, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]),
    "vx3b_model_set_material_data": (C.c_int, [_P, C.c_int, C.c_int, _DP, _DP]),
    "vx3b_job_run": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_int]),
    "vx3b_job_run_files": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_int, C.c_int]),
    "vx3b_job_run_task": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int]),
    "vx3b_job_history_files": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_int]),
    "vx3b_gen_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "vx3b_gen_set_option": (C.c_int, [_P, C.c_char_p, C.c_double]),
    "vx3b_gen_run": (C.c_int, [_P, C.POINTER(_P), C.c_int, C.c_int64, C.c_int, C.POINTER(WorldResult)]),
    "vx3b_gen_total_voxel_steps": (C.c_int64, [_P]),
    "vx3b_gen_destroy": (None, [_P]),
    "vx3b_shard_worlds": (C.c_int, [C.POINTER(C.c_int64), C.c_int, C.c_int, C.POINTER(C.c_int32)]),
}

_lib = None


def load(path=None):
    """Load libvx3b200.so and bind every declared symbol.  Raises if the library or a symbol is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(
            f"{p} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
