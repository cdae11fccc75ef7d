"""ctypes binding of libroko_b200.so (C ABI declared in include/roko_b200.h).

There is deliberately no fallback: if the library has not been built, or a call fails, this
raises.  Build with ``python -m roko_b200.build`` (or ``__graft_entry__.build()``).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libroko_b200.so")

OK, EARG, ECUDA, ESTATE, ECODES, ERANGE = 0, 1, 2, 3, 4, 5

c_model_p = ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/roko_b200.h one to one
SIGNATURES = {
    "roko_b200_abi_version": (ctypes.c_int, []),
    "roko_b200_last_error": (ctypes.c_char_p, []),
    "roko_b200_window_reads": (ctypes.c_int, []),
    "roko_b200_window_cols": (ctypes.c_int, []),
    "roko_b200_num_classes": (ctypes.c_int, []),
    "roko_b200_raw_weight_count": (ctypes.c_size_t, []),
    "roko_b200_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "roko_b200_model_create": (ctypes.c_int, [ctypes.POINTER(c_model_p), ctypes.c_int]),
    "roko_b200_model_load": (ctypes.c_int, [c_model_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]),
    "roko_b200_model_destroy": (ctypes.c_int, [c_model_p]),
    "roko_b200_forward_u8": (ctypes.c_int, [c_model_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "roko_b200_forward_i64": (ctypes.c_int, [c_model_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "roko_b200_infer_host": (ctypes.c_int, [c_model_p, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int,
                                            ctypes.c_void_p, ctypes.c_void_p]),
    "roko_b200_model_check": (ctypes.c_int, [c_model_p]),
    "roko_b200_model_set_option": (ctypes.c_int, [c_model_p, ctypes.c_char_p, ctypes.c_longlong]),
    "roko_b200_forward_taps": (ctypes.c_int, [c_model_p, ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 7
                               + [ctypes.c_size_t, ctypes.c_void_p]),
    "roko_b200_forward_timed": (ctypes.c_int, [c_model_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_int,
                                               ctypes.POINTER(ctypes.c_float)]),
    "roko_b200_measure_fp32_peak": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]),
    "roko_b200_train_workspace_bytes": (ctypes.c_size_t, [ctypes.c_int]),
    "roko_b200_train_forward": (ctypes.c_int, [c_model_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                               ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_void_p,
                                               ctypes.c_size_t, ctypes.c_void_p]),
    "roko_b200_train_backward": (ctypes.c_int, [c_model_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_float,
                                                ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_void_p,
                                                ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "roko_b200_dropout_mask": (ctypes.c_int, [ctypes.c_float, ctypes.c_ulonglong, ctypes.c_int, ctypes.c_size_t,
                                              ctypes.c_void_p, ctypes.c_void_p]),
}

_lib = None


class RokoB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"libroko_b200 error {code}: {message}")
        self.code = code


def lib():
    """Load the shared library once; raise loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the roko_b200 hot path is CUDA only and has no fallback. "
                "Build it with `python -m roko_b200.build`.")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)            # AttributeError if the .so lacks a declared symbol
            fn.restype, fn.argtypes = res, args
        if L.roko_b200_abi_version() != 1:
            raise RuntimeError("libroko_b200 ABI version mismatch; rebuild")
        _lib = L
    return _lib


def check(rc):
    if rc != OK:
        msg = lib().roko_b200_last_error()
        err = RokoB200Error(rc, msg.decode() if msg else "")
        if rc == ECODES:
            raise IndexError(str(err))       # what nn.Embedding raises for an out-of-range code
        raise err
    return rc
