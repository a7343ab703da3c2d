"""ctypes binding of ``libazp_b200.so`` -- the C ABI declared in ``include/azp_b200.h``.

The library is the product: there is no Python, PyTorch or CPU fallback for any entry point. If
the shared object is missing the import fails loudly and tells the user how to build it.
"""

import ctypes
import os

# libazp_b200.so links the shared CUDA runtime. Import torch first so that the libcudart already
# mapped by torch is the one the library binds to (one runtime instance: same current device,
# same streams).
import torch  # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AZP_B200_LIB", os.path.join(HERE, "libazp_b200.so"))  # override: A/B builds

EV_PERTURBED_LENNARD_JONES = 0
EV_EXPANDED_YUKAWA = 1
EV_COLLOID = 2
EV_HERTZ = 3
EV_DPD_GENERAL_WEIGHT = 4
EV_TWO_PATCH_MORSE = 5

FAMILY_PAIR, FAMILY_DPD, FAMILY_ANISO = 0, 1, 2
SHIFT_MODES = {"none": 0, "shift": 1, "xplor": 2}


class AzpBox(ctypes.Structure):
    _fields_ = [
        ("L", ctypes.c_double * 3),
        ("tilt", ctypes.c_double * 3),
        ("periodic", ctypes.c_int32 * 3),
        ("_pad", ctypes.c_int32),
    ]


class AzpBarrierArgs(ctypes.Structure):
    _fields_ = [
        ("d_force", ctypes.c_void_p),
        ("d_virial", ctypes.c_void_p),
        ("virial_pitch", ctypes.c_uint64),
        ("d_pos", ctypes.c_void_p),
        ("d_params", ctypes.c_void_p),
        ("box", AzpBox),
        ("location", ctypes.c_double),
        ("N", ctypes.c_uint32),
        ("ntypes", ctypes.c_uint32),
        ("geometry", ctypes.c_int32),
        ("block_size", ctypes.c_uint32),
    ]


BARRIER_PLANAR, BARRIER_SPHERICAL = 0, 1


MD_MAX_FORCES = 8


class AzpMdArgs(ctypes.Structure):
    _fields_ = [
        ("d_pos", ctypes.c_void_p),
        ("d_vel", ctypes.c_void_p),
        ("d_accel", ctypes.c_void_p),
        ("d_image", ctypes.c_void_p),
        ("d_net_force", ctypes.c_void_p),
        ("d_forces", ctypes.c_void_p * MD_MAX_FORCES),
        ("n_forces", ctypes.c_uint32),
        ("N", ctypes.c_uint32),
        ("box", AzpBox),
        ("dt", ctypes.c_double),
    ]


class AzpLangevinArgs(ctypes.Structure):
    _fields_ = [
        ("d_tag", ctypes.c_void_p),
        ("d_gamma", ctypes.c_void_p),
        ("ntypes", ctypes.c_uint32),
        ("seed", ctypes.c_uint32),
        ("timestep", ctypes.c_uint64),
        ("kT", ctypes.c_double),
        ("rng_id", ctypes.c_uint32),
        ("noiseless", ctypes.c_uint32),
    ]


class AzpWallArgs(ctypes.Structure):
    _fields_ = [
        ("d_force", ctypes.c_void_p),
        ("d_virial", ctypes.c_void_p),
        ("virial_pitch", ctypes.c_uint64),
        ("d_pos", ctypes.c_void_p),
        ("d_params", ctypes.c_void_p),
        ("d_walls", ctypes.c_void_p),
        ("N", ctypes.c_uint32),
        ("ntypes", ctypes.c_uint32),
        ("block_size", ctypes.c_uint32),
        ("_pad", ctypes.c_uint32),
    ]


WALL_COLLOID, WALL_LJ93 = 0, 1


class AzpPairArgs(ctypes.Structure):
    _fields_ = [
        ("d_force", ctypes.c_void_p),
        ("d_virial", ctypes.c_void_p),
        ("d_torque", ctypes.c_void_p),
        ("virial_pitch", ctypes.c_uint64),
        ("d_pos", ctypes.c_void_p),
        ("d_vel", ctypes.c_void_p),
        ("d_orientation", ctypes.c_void_p),
        ("d_tag", ctypes.c_void_p),
        ("d_n_neigh", ctypes.c_void_p),
        ("d_nlist", ctypes.c_void_p),
        ("d_head_list", ctypes.c_void_p),
        ("size_neigh_list", ctypes.c_uint64),
        ("d_rcutsq", ctypes.c_void_p),
        ("d_ronsq", ctypes.c_void_p),
        ("box", AzpBox),
        ("N", ctypes.c_uint32),
        ("ntypes", ctypes.c_uint32),
        ("shift_mode", ctypes.c_uint32),
        ("compute_virial", ctypes.c_uint32),
        ("block_size", ctypes.c_uint32),
        ("threads_per_particle", ctypes.c_uint32),
        ("seed", ctypes.c_uint32),
        ("n_max", ctypes.c_uint32),
        ("timestep", ctypes.c_uint64),
        ("deltaT", ctypes.c_double),
        ("T", ctypes.c_double),
        ("row_offset", ctypes.c_uint32),
        ("n_row_ids", ctypes.c_uint32),
        ("d_row_ids", ctypes.c_void_p),
    ]


class AzpNlistArgs(ctypes.Structure):
    _fields_ = [
        ("d_pos", ctypes.c_void_p),
        ("N", ctypes.c_uint32),
        ("ntypes", ctypes.c_uint32),
        ("box", AzpBox),
        ("d_rlistsq", ctypes.c_void_p),
        ("r_list_max", ctypes.c_double),
        ("d_n_neigh", ctypes.c_void_p),
        ("d_head_list", ctypes.c_void_p),
        ("d_nlist", ctypes.c_void_p),
        ("d_cell_of", ctypes.c_void_p),
        ("d_cell_start", ctypes.c_void_p),
        ("d_cell_order", ctypes.c_void_p),
        ("cell_dim", ctypes.c_uint32 * 3),
        ("row_offset", ctypes.c_uint32),
        ("n_rows", ctypes.c_uint32),
        ("_pad", ctypes.c_uint32),
        ("d_capacity", ctypes.c_void_p),
        ("d_overflow", ctypes.c_void_p),
        ("d_cell_pos", ctypes.c_void_p),
        ("d_pos_at_build", ctypes.c_void_p),
        ("threads_per_row", ctypes.c_uint32),
        ("_pad2", ctypes.c_uint32),
    ]


# every symbol include/azp_b200.h declares; tests check the library exports each one
EXPORTED_SYMBOLS = (
    "azp_abi_version",
    "azp_error_string",
    "azp_evaluator_name",
    "azp_param_num_fields",
    "azp_param_size",
    "azp_param_pack",
    "azp_param_unpack",
    "azp_pair_forces_f32",
    "azp_pair_forces_f64",
    "azp_dpd_forces_f32",
    "azp_dpd_forces_f64",
    "azp_aniso_forces_f32",
    "azp_aniso_forces_f64",
    "azp_pair_forces_fused_f32",
    "azp_pair_forces_fused_f64",
    "azp_autotune",
    "azp_gather_rows",
    "azp_push_rows",
    "azp_harmonic_barrier_f32",
    "azp_harmonic_barrier_f64",
    "azp_harmonic_barrier_valid",
    "azp_wall_forces_f32",
    "azp_wall_forces_f64",
    "azp_wall_param_size",
    "azp_walls_size",
    "azp_nve_step_one_f32",
    "azp_nve_step_one_f64",
    "azp_nve_step_two_f32",
    "azp_nve_step_two_f64",
    "azp_langevin_step_two_f32",
    "azp_langevin_step_two_f64",
    "azp_nlist_moved_f32",
    "azp_nlist_moved_f64",
    "azp_sfc_order_f32",
    "azp_sfc_order_f64",
    "azp_dpd_alpha",
    "azp_philox4x32_10",
    "azp_nlist_cell_dim",
    "azp_nlist_bin_f32",
    "azp_nlist_bin_f64",
    "azp_nlist_count_f32",
    "azp_nlist_count_f64",
    "azp_nlist_fill_f32",
    "azp_nlist_fill_f64",
)


class AzpError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "azplugins_b200: %s is missing. The CUDA extension is the only implementation of the "
            "pair-force path (no CPU fallback). Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C azplugins_b200/csrc`." % LIB_PATH
        )
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, u32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32
    lib.azp_abi_version.restype = i32
    lib.azp_error_string.argtypes = [i32]
    lib.azp_error_string.restype = ctypes.c_char_p
    lib.azp_evaluator_name.argtypes = [i32]
    lib.azp_evaluator_name.restype = ctypes.c_char_p
    lib.azp_param_num_fields.argtypes = [i32]
    lib.azp_param_size.argtypes = [i32, i32]
    lib.azp_param_pack.argtypes = [i32, i32, vp, vp]
    lib.azp_param_unpack.argtypes = [i32, i32, vp, vp]
    for name in ("azp_pair_forces_f32", "azp_pair_forces_f64", "azp_dpd_forces_f32",
                 "azp_dpd_forces_f64"):
        getattr(lib, name).argtypes = [i32, ctypes.POINTER(AzpPairArgs), vp, vp]
        getattr(lib, name).restype = i32
    for name in ("azp_aniso_forces_f32", "azp_aniso_forces_f64"):
        getattr(lib, name).argtypes = [i32, ctypes.POINTER(AzpPairArgs), vp, vp, vp]
        getattr(lib, name).restype = i32
    for name in ("azp_pair_forces_fused_f32", "azp_pair_forces_fused_f64"):
        if hasattr(lib, name):  # absent from pre-fusion A/B builds loaded through AZP_B200_LIB
            getattr(lib, name).argtypes = [i32, ctypes.POINTER(AzpPairArgs), vp, i32,
                                           ctypes.POINTER(AzpPairArgs), vp, vp]
            getattr(lib, name).restype = i32
    lib.azp_autotune.argtypes = [i32, i32, i32, ctypes.POINTER(AzpPairArgs), vp, vp,
                                 ctypes.POINTER(u32), ctypes.POINTER(u32),
                                 ctypes.POINTER(ctypes.c_float)]
    lib.azp_gather_rows.argtypes = [vp, vp, ctypes.c_uint64, u32, vp, vp]
    lib.azp_gather_rows.restype = i32
    lib.azp_push_rows.argtypes = [vp, vp, vp, ctypes.c_uint64, u32, vp]
    lib.azp_push_rows.restype = i32
    for sfx in ("_f32", "_f64"):
        fn = getattr(lib, "azp_harmonic_barrier" + sfx)
        fn.argtypes = [ctypes.POINTER(AzpBarrierArgs), vp]
        fn.restype = i32
    lib.azp_harmonic_barrier_valid.argtypes = [i32, i32, ctypes.c_double, ctypes.POINTER(AzpBox)]
    lib.azp_harmonic_barrier_valid.restype = i32
    for sfx in ("_f32", "_f64"):
        fn = getattr(lib, "azp_wall_forces" + sfx)
        fn.argtypes = [i32, ctypes.POINTER(AzpWallArgs), vp]
        fn.restype = i32
    lib.azp_wall_param_size.argtypes = [i32, i32]
    lib.azp_wall_param_size.restype = i32
    lib.azp_walls_size.argtypes = [i32]
    lib.azp_walls_size.restype = i32
    for name in ("azp_nve_step_one", "azp_nve_step_two"):
        for sfx in ("_f32", "_f64"):
            fn = getattr(lib, name + sfx)
            fn.argtypes = [ctypes.POINTER(AzpMdArgs), vp]
            fn.restype = i32
    for sfx in ("_f32", "_f64"):
        fn = getattr(lib, "azp_langevin_step_two" + sfx)
        fn.argtypes = [ctypes.POINTER(AzpMdArgs), ctypes.POINTER(AzpLangevinArgs), vp]
        fn.restype = i32
    lib.azp_dpd_alpha.argtypes = [i32, u32, u32, u32, ctypes.c_uint64]
    lib.azp_dpd_alpha.restype = ctypes.c_double
    lib.azp_philox4x32_10.argtypes = [vp, vp, vp]
    lib.azp_philox4x32_10.restype = None
    lib.azp_nlist_cell_dim.argtypes = [ctypes.POINTER(AzpBox), ctypes.c_double, vp]
    for name in ("azp_nlist_bin_f32", "azp_nlist_bin_f64", "azp_nlist_count_f32",
                 "azp_nlist_count_f64", "azp_nlist_fill_f32", "azp_nlist_fill_f64"):
        getattr(lib, name).argtypes = [ctypes.POINTER(AzpNlistArgs), vp]
        getattr(lib, name).restype = i32
    for sfx in ("_f32", "_f64"):
        fn = getattr(lib, "azp_nlist_moved" + sfx)
        fn.argtypes = [vp, vp, ctypes.POINTER(AzpBox), ctypes.c_double, u32, vp, vp]
        fn.restype = i32
        fn = getattr(lib, "azp_sfc_order" + sfx)
        fn.argtypes = [vp, ctypes.POINTER(AzpBox), u32, vp, vp]
        fn.restype = i32
    if lib.azp_abi_version() != 1:
        raise ImportError("azplugins_b200: ABI version mismatch in " + LIB_PATH)
    return lib


lib = _load()


def check(code, what=""):
    if code != 0:
        msg = lib.azp_error_string(code)
        raise AzpError("%s failed: %s (code %d)" % (what or "azp call", msg.decode() if msg else "?", code))
