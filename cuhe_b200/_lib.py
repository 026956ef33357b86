"""ctypes binding of include/cuhe_b200.h (libcuhe_b200.so, built in-tree by
__graft_entry__.build()).  There is NO fallback: if the CUDA library is missing
or fails to load, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CUHE_B200_LIB", os.path.join(_HERE, "libcuhe_b200.so"))   # override: A/B kernel experiments


class CuHEError(RuntimeError):
    """Raised where the reference prints a message and calls terminate()/exit()
    (cuhe/CuHE.cu:102-113, cuhe/Debug.h:39-53)."""


class cuhe_params(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "mSize", "modLen", "modLen2", "rawLen", "crtLen", "nttLen",
        "logCoeffMax", "logCoeffMin", "logCoeffCut",
        "depth", "modMsg", "logMsg", "wordsMsg",
        "logRelin", "numEvalKey", "logCrtPrime", "numCrtPrime")]


# every symbol include/cuhe_b200.h declares: name -> (restype, argtypes)
_vp, _i, _ll = C.c_void_p, C.c_int, C.c_longlong
_pp = C.POINTER(cuhe_params)
SYMBOLS = {
    "cuhe_version": (_i, []),
    "cuhe_last_error": (C.c_char_p, []),
    "cuhe_set_parameters": (_i, [_pp, _i, _i, _i, _i, _i, _i]),
    "cuhe_param_num_crt_prime": (_i, [_pp, _i]),
    "cuhe_param_log_coeff": (_i, [_pp, _i]),
    "cuhe_param_words_coeff": (_i, [_pp, _i]),
    "cuhe_param_num_eval_key": (_i, [_pp, _i]),
    "cuhe_param_get_level": (_i, [_pp, _i]),
    "cuhe_ctx_create": (_i, [C.POINTER(_vp), _pp, _i, _i, _i]),
    "cuhe_ctx_destroy": (_i, [_vp]),
    "cuhe_ctx_params": (_i, [_vp, _pp]),
    "cuhe_ctx_crt_primes_host": (_i, [_vp, _vp]),
    "cuhe_ctx_coeff_modulus_host": (_i, [_vp, _i, _vp, _i]),
    "cuhe_ctx_rows": (_i, [_vp, _i]),
    "cuhe_ctx_set_poly_modulus_host": (_i, [_vp, _vp, _i]),
    "cuhe_malloc": (_i, [_vp, C.POINTER(_vp), C.c_size_t, _vp]),
    "cuhe_free": (_i, [_vp, _vp, _vp]),
    "cuhe_pool_trim": (_i, [_vp]),
    "cuhe_memcpy": (_i, [_vp, _vp, _vp, C.c_size_t, _i, _vp]),
    "cuhe_memset": (_i, [_vp, _vp, _i, C.c_size_t, _vp]),
    "cuhe_stream_sync": (_i, [_vp, _vp]),
    "cuhe_host_alloc": (_i, [C.POINTER(_vp), C.c_size_t]),
    "cuhe_host_free": (_i, [_vp]),
    "cuhe_device_count": (_i, []),
    "cuhe_crt": (_i, [_vp, _vp, _vp, _i, _vp]),
    "cuhe_icrt": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "cuhe_ntt": (_i, [_vp, _vp, _vp, _i, _vp]),
    "cuhe_intt": (_i, [_vp, _vp, _vp, _i, _vp]),
    "cuhe_intt_double_deg": (_i, [_vp, _vp, _vp, _i, _vp]),
    "cuhe_intt_mod": (_i, [_vp, _vp, _vp, _i, _vp]),
    "cuhe_barrett": (_i, [_vp, _vp, _vp, _i, _vp]),
    "cuhe_ntt_mul": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "cuhe_ntt_add": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "cuhe_ntt_mul_nx1": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "cuhe_ntt_add_nx1": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "cuhe_ntt_mul_intt_mod": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "cuhe_crt_add": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "cuhe_crt_add_int": (_i, [_vp, _vp, _vp, C.c_uint, _i, _vp]),
    "cuhe_crt_add_nx1": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "cuhe_mod_switch": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "cuhe_relin_init": (_i, [_vp, _vp, _vp]),
    "cuhe_relin": (_i, [_vp, _vp, _vp, _i, _vp]),
    "cuhe_relin_key_words": (C.c_size_t, [_vp]),
    "cuhe_relin_export_host": (_i, [_vp, _vp, C.c_size_t, _vp]),
    "cuhe_relin_import_host": (_i, [_vp, _vp, C.c_size_t, _vp]),
    "cuhe_ntt_ext_batch": (_i, [_vp, _vp, _vp, _i, _i, _ll, _vp]),
    "cuhe_intt_batch": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "cuhe_mul_raw_host": (_i, [_vp, _vp, _vp, _vp, _i, _vp]),
    "cuhe_mul_raw_host_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "cuhe_mul_crt_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "cuhe_icrt_batch": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "cuhe_icrt_slice_batch": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "cuhe_crt_batch": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "cuhe_ntt_batch": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "cuhe_ntt_mul_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "cuhe_intt_mod_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "cuhe_crt_add_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "cuhe_crt_add_int_batch": (_i, [_vp, _vp, _vp, C.c_uint, _i, _i, _vp]),
    "cuhe_mod_switch_batch": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "cuhe_relin_batch": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "cuhe_comm_unique_id": (_i, [_vp]),
    "cuhe_ctx_comm_init": (_i, [_vp, _vp]),
    "cuhe_ctx_comm_attach": (_i, [_vp, _vp]),
    "cuhe_mul_raw_sharded_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp]),
    "cuhe_modp_batch": (_i, [_vp, _i, _vp, _vp, _vp, C.c_size_t, _i, _vp]),
    "cuhe_launch_count": (_ll, [_i]),
}

_lib = None


def load_library():
    """Load libcuhe_b200.so and bind every declared symbol; raises if the
    library is absent (no CPU path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CuHEError(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)      # AttributeError if the export is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load_library().cuhe_last_error().decode(errors="replace")
        raise CuHEError(f"cuhe error {rc}: {msg}")
