"""ctypes binding of libmico_b200.so (the C-ABI in include/mico_b200.h).

The library is REQUIRED: there is no CPU or PyTorch fallback.  Importing this module when the
.so is missing raises, and every op raises if the call returns a non-zero MICO_ERR_* code.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MICO_B200_LIB selects another build of the same library (A/B measurements of a kernel change on one box)
LIB_PATH = os.environ.get("MICO_B200_LIB") or os.path.join(_HERE, "lib", "libmico_b200.so")

ACT_NONE, ACT_GELU, ACT_QUICK_GELU, ACT_GELU_BWD, ACT_QUICK_GELU_BWD = 0, 1, 2, 3, 4
ACT_GELU_SAVE_GRAD, ACT_QUICK_GELU_SAVE_GRAD, ACT_MUL_AUX = 5, 6, 7


class MicoError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", C.c_void_p), ("lda", C.c_int64), ("a_mn_major", C.c_int32),
        ("b", C.c_void_p), ("ldb", C.c_int64), ("b_mn_major", C.c_int32),
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("out", C.c_void_p), ("ldo", C.c_int64), ("out_fp32", C.c_int32),
        ("bias", C.c_void_p),
        ("residual", C.c_void_p), ("ldr", C.c_int64),
        ("row_scale", C.c_void_p), ("rows_per_group", C.c_int32),
        ("act", C.c_int32),
        ("aux_out", C.c_void_p), ("ld_aux_out", C.c_int64),
        ("aux_in", C.c_void_p), ("ld_aux_in", C.c_int64),
        ("accumulate", C.c_int32),
        ("alpha", C.c_float),
        ("remap_gin", C.c_int32), ("remap_gout", C.c_int32), ("remap_off", C.c_int32),
        ("residual_bcast", C.c_int32),
        ("asum_out", C.c_void_p), ("ones", C.c_void_p),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise MicoError(
            f"{LIB_PATH} not found: build it with `python -m mico_b200.build` "
            "(mico_b200 has no CPU/PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    lib.mico_version.restype = C.c_int
    lib.mico_last_error.restype = C.c_char_p
    lib.mico_launch_count.restype = C.c_int64
    lib.mico_reset_launch_count.restype = None
    return lib


lib = _load()


def check(rc, what):
    if rc != 0:
        msg = lib.mico_last_error().decode("utf-8", "replace")
        raise MicoError(f"{what} failed with code {rc}: {msg}")


def launch_count():
    return int(lib.mico_launch_count())


def reset_launch_count():
    lib.mico_reset_launch_count()


class AttnArgs(C.Structure):
    _fields_ = [
        ("q", C.c_void_p), ("q_bs", C.c_int64), ("q_rs", C.c_int64), ("q_hs", C.c_int64),
        ("k", C.c_void_p), ("k_bs", C.c_int64), ("k_rs", C.c_int64), ("k_hs", C.c_int64),
        ("v", C.c_void_p), ("v_bs", C.c_int64), ("v_rs", C.c_int64), ("v_hs", C.c_int64),
        ("o", C.c_void_p), ("o_bs", C.c_int64), ("o_rs", C.c_int64), ("o_hs", C.c_int64),
        ("lse", C.c_void_p),
        ("mask", C.c_void_p), ("mask_bs", C.c_int64), ("mask_qs", C.c_int64),
        ("B", C.c_int32), ("H", C.c_int32), ("Sq", C.c_int32), ("Sk", C.c_int32), ("D", C.c_int32),
        ("scale", C.c_float),
        ("dout", C.c_void_p), ("do_bs", C.c_int64), ("do_rs", C.c_int64), ("do_hs", C.c_int64),
        ("delta", C.c_void_p),
        ("dq", C.c_void_p), ("dq_bs", C.c_int64), ("dq_rs", C.c_int64), ("dq_hs", C.c_int64),
        ("dk", C.c_void_p), ("dk_bs", C.c_int64), ("dk_rs", C.c_int64), ("dk_hs", C.c_int64),
        ("dv", C.c_void_p), ("dv_bs", C.c_int64), ("dv_rs", C.c_int64), ("dv_hs", C.c_int64),
        ("mask_hs", C.c_int64), ("mask_bmod", C.c_int32),
        ("dropout_p", C.c_float), ("dropout_seed", C.c_uint64),
        ("kv_index", C.c_void_p), ("n_kv", C.c_int32), ("grp_ptr", C.c_void_p), ("grp_list", C.c_void_p),
    ]


lib.mico_layernorm_bwd_workspace.restype = C.c_size_t
lib.mico_colsum_workspace.restype = C.c_size_t
