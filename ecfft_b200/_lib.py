"""ctypes loader for ecfft_b200/lib/libecfft_b200.so (the C ABI declared in include/ecfft_b200.h).

There is no CPU fallback: if the CUDA library is missing, or a call needs a GPU that is not
there, the failure is loud (ImportError / EcfftError).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ECFFT_B200_LIB") or os.path.join(_HERE, "lib", "libecfft_b200.so")  # override: A/B experiments only

ECFFT_OK = 0
ERR_NOT_POW2 = 1
ERR_TREE_TOO_SMALL = 2
ERR_BAD_BYTES = 3
ERR_CUDA = 4
ERR_INVALID_ARG = 5
ERR_TOO_LARGE = 6
ERR_MISSING_TABLES = 7
ERR_BUFFER_TOO_SMALL = 8

# every symbol include/ecfft_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "ecfft_last_error", "ecfft_device_count",
    "ecfft_tree_build_secp256k1", "ecfft_tree_new", "ecfft_tree_deserialize",
    "ecfft_tree_serialized_size", "ecfft_tree_serialize", "ecfft_tree_free",
    "ecfft_tree_leaves", "ecfft_tree_device", "ecfft_tree_table",
    "ecfft_enter", "ecfft_exit", "ecfft_extend", "ecfft_mextend", "ecfft_degree",
    "ecfft_redc_z0", "ecfft_redc_z1", "ecfft_modular_reduce", "ecfft_vanish",
    "ecfft_enter_dev", "ecfft_exit_dev", "ecfft_extend_dev", "ecfft_mextend_dev",
    "ecfft_degree_dev", "ecfft_redc_z0_dev", "ecfft_redc_z1_dev",
    "ecfft_modular_reduce_dev", "ecfft_vanish_dev", "ecfft_enter_range_dev",
    "ecfft_launch_count", "ecfft_profile_enable", "ecfft_profile_read",
    "ecfft_mg_prescale_dev", "ecfft_mg_cross_dev", "ecfft_mg_local_dev", "ecfft_mg_combine_dev",
    "ecfft_mg_arena_alloc", "ecfft_mg_arena_open", "ecfft_mg_arena_close", "ecfft_mg_arena_free", "ecfft_mg_arena_reset", "ecfft_mg_arena_status",
    "ecfft_mg_signal_dev", "ecfft_mg_wait_dev", "ecfft_mg_arena_bytes", "ecfft_enter_peer_dev",
    "ecfft_selftest_field", "ecfft_flow_stats", "ecfft_pointwise_mul", "ecfft_pointwise_mul_dev",
    "ecfft_mg_exit_arena_bytes", "ecfft_exit_peer_dev", "ecfft_enter_many",
    "ecfft_m31_tree_build", "ecfft_m31_tree_free", "ecfft_m31_tree_leaves", "ecfft_m31_tree_table",
    "ecfft_m31_enter", "ecfft_m31_exit", "ecfft_m31_extend", "ecfft_m31_mextend", "ecfft_m31_degree",
    "ecfft_m31_redc_z0", "ecfft_m31_redc_z1", "ecfft_m31_modular_reduce", "ecfft_m31_vanish",
    "ecfft_m31_enter_dev", "ecfft_m31_exit_dev", "ecfft_m31_extend_dev",
]


class EcfftError(RuntimeError):
    """Raised where the reference would panic or return a SerializationError."""

    def __init__(self, code, message):
        super().__init__(f"ecfft_b200 error {code}: {message}")
        self.code = code


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C ecfft_b200/csrc`). ecfft_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    psz, pvp = ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_void_p)
    L.ecfft_last_error.restype = ctypes.c_char_p
    L.ecfft_last_error.argtypes = []
    L.ecfft_device_count.argtypes = [ctypes.POINTER(ci)]
    L.ecfft_tree_build_secp256k1.argtypes = [sz, ci, ci, pvp]
    L.ecfft_tree_new.argtypes = [vp, sz, vp, vp, sz, ci, ci, pvp]
    L.ecfft_tree_deserialize.argtypes = [vp, sz, ci, ci, pvp]
    L.ecfft_tree_serialized_size.argtypes = [vp, ci, psz]
    L.ecfft_tree_serialize.argtypes = [vp, ci, vp, sz, psz]
    L.ecfft_tree_free.argtypes = [vp]
    L.ecfft_tree_free.restype = None
    L.ecfft_tree_leaves.argtypes = [vp]
    L.ecfft_tree_leaves.restype = sz
    L.ecfft_tree_device.argtypes = [vp]
    L.ecfft_tree_table.argtypes = [vp, sz, ctypes.c_char_p, vp, sz, psz]
    for name in ("ecfft_enter", "ecfft_exit"):
        getattr(L, name).argtypes = [vp, vp, sz, vp]
    for name in ("ecfft_extend", "ecfft_mextend"):
        getattr(L, name).argtypes = [vp, vp, sz, ci, vp]
    L.ecfft_degree.argtypes = [vp, vp, sz, psz]
    L.ecfft_enter_many.argtypes = [vp, vp, sz, sz, vp]
    for name in ("ecfft_redc_z0", "ecfft_redc_z1"):
        getattr(L, name).argtypes = [vp, vp, vp, sz, vp]
    L.ecfft_modular_reduce.argtypes = [vp, vp, vp, vp, sz, vp]
    L.ecfft_pointwise_mul.argtypes = [vp, vp, vp, sz, vp]
    L.ecfft_pointwise_mul_dev.argtypes = [vp, vp, vp, sz, vp, vp]
    L.ecfft_vanish.argtypes = [vp, vp, sz, vp]
    for name in ("ecfft_enter_dev", "ecfft_exit_dev"):
        getattr(L, name).argtypes = [vp, vp, sz, vp, vp]
    for name in ("ecfft_extend_dev", "ecfft_mextend_dev"):
        getattr(L, name).argtypes = [vp, vp, sz, ci, vp, vp]
    L.ecfft_degree_dev.argtypes = [vp, vp, sz, psz, vp]
    for name in ("ecfft_redc_z0_dev", "ecfft_redc_z1_dev"):
        getattr(L, name).argtypes = [vp, vp, vp, sz, vp, vp]
    L.ecfft_modular_reduce_dev.argtypes = [vp, vp, vp, vp, sz, vp, vp]
    L.ecfft_vanish_dev.argtypes = [vp, vp, sz, vp, vp]
    L.ecfft_enter_range_dev.argtypes = [vp, vp, sz, sz, sz, vp, vp]
    L.ecfft_mg_prescale_dev.argtypes = [vp, sz, sz, vp, sz, vp, vp]
    L.ecfft_mg_cross_dev.argtypes = [vp, sz, ci, ctypes.c_uint, ci, sz, vp, vp, sz, vp, vp]
    L.ecfft_mg_local_dev.argtypes = [vp, sz, vp, sz, vp, vp]
    L.ecfft_mg_combine_dev.argtypes = [vp, sz, sz, vp, vp, vp, vp, sz, vp, vp]
    L.ecfft_mg_arena_alloc.argtypes = [ci, sz, ctypes.POINTER(vp), ctypes.c_char_p]
    L.ecfft_mg_arena_open.argtypes = [ci, ctypes.c_char_p, ctypes.POINTER(vp)]
    L.ecfft_mg_arena_close.argtypes = [vp]
    L.ecfft_mg_arena_free.argtypes = [vp]
    L.ecfft_mg_arena_reset.argtypes = [vp, vp]
    L.ecfft_mg_arena_status.argtypes = [vp, ctypes.POINTER(ctypes.c_ulonglong)]
    L.ecfft_selftest_field.argtypes = [ci, ctypes.c_ulonglong, ctypes.POINTER(ctypes.c_ulonglong)]
    L.ecfft_mg_arena_bytes.argtypes = [sz, ci, psz]
    L.ecfft_enter_peer_dev.argtypes = [vp, vp, sz, ci, ci, pvp, ctypes.c_ulonglong, vp, vp]
    L.ecfft_exit_peer_dev.argtypes = [vp, vp, sz, ci, ci, pvp, ctypes.c_ulonglong, vp, vp]
    L.ecfft_mg_exit_arena_bytes.argtypes = [sz, ci, psz]
    L.ecfft_mg_signal_dev.argtypes = [vp, ctypes.c_ulonglong, vp]
    L.ecfft_mg_wait_dev.argtypes = [vp, ctypes.c_ulonglong, ctypes.c_uint, vp]
    L.ecfft_flow_stats.argtypes = [ci, ctypes.POINTER(ctypes.c_ulonglong)]
    L.ecfft_m31_tree_build.argtypes = [sz, ci, pvp]
    L.ecfft_m31_tree_free.argtypes = [vp]
    L.ecfft_m31_tree_free.restype = None
    L.ecfft_m31_tree_leaves.argtypes = [vp]
    L.ecfft_m31_tree_leaves.restype = sz
    L.ecfft_m31_tree_table.argtypes = [vp, sz, ctypes.c_char_p, vp, sz, psz]
    for name in ("ecfft_m31_enter", "ecfft_m31_exit", "ecfft_m31_vanish"):
        getattr(L, name).argtypes = [vp, vp, sz, vp]
    for name in ("ecfft_m31_extend", "ecfft_m31_mextend"):
        getattr(L, name).argtypes = [vp, vp, sz, ci, vp]
    L.ecfft_m31_degree.argtypes = [vp, vp, sz, psz]
    for name in ("ecfft_m31_redc_z0", "ecfft_m31_redc_z1"):
        getattr(L, name).argtypes = [vp, vp, vp, sz, vp]
    L.ecfft_m31_modular_reduce.argtypes = [vp, vp, vp, vp, sz, vp]
    for name in ("ecfft_m31_enter_dev", "ecfft_m31_exit_dev"):
        getattr(L, name).argtypes = [vp, vp, sz, vp, vp]
    L.ecfft_m31_extend_dev.argtypes = [vp, vp, sz, ci, vp, vp]
    L.ecfft_launch_count.restype = ctypes.c_ulonglong
    L.ecfft_launch_count.argtypes = []
    L.ecfft_profile_enable.restype = None
    L.ecfft_profile_enable.argtypes = [ci]
    L.ecfft_profile_read.argtypes = [ci, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                     ctypes.POINTER(ctypes.c_ulonglong)]
    _lib = L
    return L


def check(rc):
    if rc != ECFFT_OK:
        raise EcfftError(rc, load().ecfft_last_error().decode(errors="replace"))
