"""ctypes binding of include/sbk.h -- one-to-one, no logic."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_int64_p = ctypes.POINTER(ctypes.c_int64)


class SbkError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("sbk error %d: %s" % (code, message))
        self.code = code


class BodyDesc(ctypes.Structure):
    _fields_ = [("parent", ctypes.c_int32), ("joint_type", ctypes.c_int32), ("mass", ctypes.c_double),
                ("com_B", ctypes.c_double * 3), ("unit_inertia_OB_B", ctypes.c_double * 6),
                ("X_PF", ctypes.c_double * 12), ("X_BM", ctypes.c_double * 12)]


class ForceDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("body", ctypes.c_int32), ("coord", ctypes.c_int32), ("pad_", ctypes.c_int32),
                ("a", ctypes.c_double), ("b", ctypes.c_double), ("dir", ctypes.c_double * 3), ("station2", ctypes.c_double * 3)]


class RkmOpts(ctypes.Structure):
    _fields_ = [("accuracy", ctypes.c_double), ("constraint_tol", ctypes.c_double),
                ("use_infinity_norm", ctypes.c_int32), ("project_every_step", ctypes.c_int32)]


class AdaptiveOpts(ctypes.Structure):
    _fields_ = [("accuracy", ctypes.c_double), ("constraint_tol", ctypes.c_double), ("init_step", ctypes.c_double),
                ("min_step", ctypes.c_double), ("max_step", ctypes.c_double), ("use_infinity_norm", ctypes.c_int32),
                ("project_every_step", ctypes.c_int32), ("allow_interpolation", ctypes.c_int32), ("max_attempts", ctypes.c_int32)]


# every symbol include/sbk.h declares: name -> (restype, argtypes)
_P = ctypes.c_void_p
SYMBOLS = {
    "sbk_version": (ctypes.c_int, []),
    "sbk_last_error": (ctypes.c_char_p, []),
    "sbk_device_count": (ctypes.c_int, []),
    "sbk_topology_create": (_P, [ctypes.POINTER(BodyDesc), ctypes.c_int, ctypes.POINTER(ForceDesc), ctypes.c_int]),
    "sbk_topology_create_ex": (_P, [ctypes.POINTER(BodyDesc), ctypes.c_int, ctypes.POINTER(ForceDesc), ctypes.c_int, ctypes.c_uint]),
    "sbk_topology_destroy": (None, [_P]),
    "sbk_topology_counts": (ctypes.c_int, [_P, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "sbk_topology_slots": (ctypes.c_int, [_P, c_int_p, c_int_p, c_int_p, c_int_p, c_int_p]),
    "sbk_model_text": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int]),
    "sbk_topology_from_text": (_P, [ctypes.c_char_p]),
    "sbk_batch_create": (_P, [_P, ctypes.c_int, ctypes.c_int, _P]),
    "sbk_batch_destroy": (None, [_P]),
    "sbk_batch_size": (ctypes.c_int, [_P]),
    "sbk_batch_set_plan": (ctypes.c_int, [_P, ctypes.c_int]),
    "sbk_batch_get_plan": (ctypes.c_int, [_P]),
    "sbk_synchronize": (ctypes.c_int, [_P]),
    "sbk_set_state": (ctypes.c_int, [_P, c_double_p, c_double_p, c_double_p]),
    "sbk_get_state": (ctypes.c_int, [_P, c_double_p, c_double_p, c_double_p]),
    "sbk_set_state_async": (ctypes.c_int, [_P, c_double_p, c_double_p, c_double_p]),
    "sbk_get_state_async": (ctypes.c_int, [_P, c_double_p, c_double_p, c_double_p]),
    "sbk_integrator_kernel_name": (ctypes.c_int, [_P, ctypes.c_char_p, ctypes.c_int]),
    "sbk_set_state_aos": (ctypes.c_int, [_P, c_double_p, c_double_p]),
    "sbk_get_state_aos": (ctypes.c_int, [_P, c_double_p, c_double_p]),
    "sbk_state_device_ptrs": (ctypes.c_int, [_P, ctypes.POINTER(_P), ctypes.POINTER(_P), ctypes.POINTER(_P)]),
    "sbk_state_touched": (ctypes.c_int, [_P]),
    "sbk_realize_position": (ctypes.c_int, [_P]),
    "sbk_realize_velocity": (ctypes.c_int, [_P]),
    "sbk_realize_articulated_body_inertias": (ctypes.c_int, [_P]),
    "sbk_realize_acceleration": (ctypes.c_int, [_P]),
    "sbk_get_udot": (ctypes.c_int, [_P, c_double_p]),
    "sbk_get_qdot": (ctypes.c_int, [_P, c_double_p]),
    "sbk_get_qdotdot": (ctypes.c_int, [_P, c_double_p]),
    "sbk_get_qerr": (ctypes.c_int, [_P, c_double_p]),
    "sbk_get_body_transforms": (ctypes.c_int, [_P, c_double_p]),
    "sbk_get_body_velocities": (ctypes.c_int, [_P, c_double_p]),
    "sbk_get_body_accelerations": (ctypes.c_int, [_P, c_double_p]),
    "sbk_get_applied_forces": (ctypes.c_int, [_P, c_double_p, c_double_p]),
    "sbk_calc_energy": (ctypes.c_int, [_P, c_double_p, c_double_p]),
    "sbk_calc_mobilizer_reaction_forces": (ctypes.c_int, [_P, c_double_p]),
    "sbk_calc_composite_body_inertias": (ctypes.c_int, [_P, c_double_p]),
    "sbk_multiply_by_system_jacobian": (ctypes.c_int, [_P, c_double_p, c_double_p]),
    "sbk_multiply_by_system_jacobian_transpose": (ctypes.c_int, [_P, c_double_p, c_double_p]),
    "sbk_calc_acceleration": (ctypes.c_int, [_P, c_double_p, c_double_p, c_double_p, c_double_p]),
    "sbk_multiply_by_M": (ctypes.c_int, [_P, c_double_p, c_double_p]),
    "sbk_multiply_by_MInv": (ctypes.c_int, [_P, c_double_p, c_double_p]),
    "sbk_calc_residual_force": (ctypes.c_int, [_P, c_double_p, c_double_p, c_double_p, c_double_p]),
    "sbk_rkm_default_opts": (None, [ctypes.POINTER(RkmOpts)]),
    "sbk_rkm_step": (ctypes.c_int, [_P, ctypes.c_double, ctypes.c_int, ctypes.POINTER(RkmOpts), c_double_p]),
    "sbk_adaptive_default_opts": (None, [ctypes.POINTER(AdaptiveOpts)]),
    "sbk_rkm_adaptive": (ctypes.c_int, [_P, ctypes.c_double, ctypes.POINTER(AdaptiveOpts), ctypes.POINTER(ctypes.c_int32),
                                        ctypes.POINTER(ctypes.c_int32), c_double_p]),
    "sbk_rkm_stats": (ctypes.c_int, [_P, c_int64_p, c_int64_p, c_int64_p]),
    "sbk_get_status": (ctypes.c_int, [_P, ctypes.POINTER(ctypes.c_int32), c_int64_p]),
    "sbk_launch_count": (ctypes.c_int64, [_P]),
    "sbk_last_kernel_ms": (ctypes.c_double, [_P]),
    "sbk_mem_pattern_probe": (ctypes.c_int, [ctypes.c_int] * 7 + [c_double_p]),
    "sbk_dfma_probe": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_double_p]),
}


def library_path():
    # SBK_LIB lets experiments load an alternative build of the same ABI (kernel tuning variants)
    return os.environ.get("SBK_LIB") or os.path.join(_HERE, "libsbk.so")


def load_library():
    """Load libsbk.so (built in-tree by __graft_entry__.build / csrc/Makefile).  Fails loudly."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise SbkError(-1, "%s is missing: build it with `make -C simbody_b200/csrc` (there is no fallback)" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(lib, rc):
    if rc != 0:
        raise SbkError(rc, lib.sbk_last_error().decode(errors="replace"))
