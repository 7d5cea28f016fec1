"""ctypes binding of the C ABI declared in include/esr_b200.h (libesr_b200.so).

The library is built in-tree by `__graft_entry__.build()` / `python ntire2022_esr_b200/build.py`.
There is deliberately no fallback: if the shared object is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libesr_b200.so")

ARCH_IMDN, ARCH_RFDN, ARCH_RLFN, ARCH_BSRN, ARCH_RFDN_PRUNED, ARCH_FMEN = 0, 1, 2, 3, 4, 5
DTYPE_F32, DTYPE_F16 = 0, 1
OK, E_INVALID, E_STATE, E_WEIGHTS, E_CUDA, E_NOGPU = 0, -1, -2, -3, -4, -5

# every symbol include/esr_b200.h declares: (name, restype, argtypes)
_c = ctypes
SYMBOLS = [
    ("esr_create", _c.c_int, [_c.POINTER(_c.c_void_p), _c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    ("esr_load_weights", _c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_void_p, _c.POINTER(_c.c_int64), _c.c_int]),
    ("esr_finalize", _c.c_int, [_c.c_void_p]),
    ("esr_workspace_bytes", _c.c_size_t, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    ("esr_forward", _c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                               _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    ("esr_workspace_bytes_u8", _c.c_size_t, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    ("esr_forward_u8", _c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_float, _c.c_int,
                                  _c.c_void_p, _c.c_size_t, _c.c_void_p]),
    ("esr_forward_host_u8", _c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_float,
                                       _c.c_int]),
    ("esr_forward_host_u8_async", _c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_float,
                                             _c.c_int, _c.POINTER(_c.c_longlong)]),
    ("esr_forward_host", _c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    ("esr_forward_host_async", _c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                          _c.POINTER(_c.c_longlong)]),
    ("esr_host_wait", _c.c_int, [_c.c_void_p, _c.c_longlong]),
    ("esr_launch_count", _c.c_int, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    ("esr_launch_name", _c.c_char_p, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    ("esr_launch_flops", _c.c_double, [_c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int, _c.c_int]),
    ("esr_profile_launches", _c.c_int, [_c.c_void_p, _c.c_void_p, _c.c_void_p, _c.c_int, _c.c_int, _c.c_int, _c.c_int,
                                        _c.c_void_p, _c.c_size_t, _c.c_int, _c.POINTER(_c.c_float), _c.c_int, _c.c_void_p]),
    ("esr_set_option", _c.c_int, [_c.c_void_p, _c.c_char_p, _c.c_int]),
    ("esr_debug_timeline", _c.c_int, [_c.c_void_p, _c.POINTER(_c.c_longlong), _c.c_int]),
    ("esr_debug_tc_layer", _c.c_int, [_c.c_void_p, _c.c_int, _c.c_char_p, _c.c_int, _c.POINTER(_c.c_int32),
                                      _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int32), _c.POINTER(_c.c_float),
                                      _c.POINTER(_c.c_float), _c.c_void_p, _c.c_size_t]),
    ("esr_debug_chain", _c.c_int, [_c.c_void_p, _c.c_int, _c.POINTER(_c.c_int32), _c.POINTER(_c.c_int32), _c.c_void_p, _c.c_size_t]),
    ("esr_last_error", _c.c_char_p, [_c.c_void_p]),
    ("esr_destroy", None, [_c.c_void_p]),
    ("esr_version", _c.c_char_p, []),
    ("esr_device_ok", _c.c_int, [_c.c_int]),
]


def load_library(path: str = LIB_PATH) -> ctypes.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build the CUDA extension first (python ntire2022_esr_b200/build.py); "
            "this engine has no CPU / PyTorch fallback")
    lib = ctypes.CDLL(path)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = load_library()


class EsrError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"esr_b200 error {code}: {message}")
        self.code = code
