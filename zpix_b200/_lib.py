"""ctypes binding of libzpixcuda.so (include/zpix_cuda.h).

The library is the product: if it is missing this module raises at import time
with the build command -- there is no Python or CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("ZPX_LIB_PATH") or os.path.join(_HERE, "libzpixcuda.so")  # (override: kernel experiments)


class ZpxImageInfo(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("width", C.c_int32), ("height", C.c_int32), ("num_components", C.c_int32),
        ("variant", C.c_int32), ("subsample_ratio", C.c_int32), ("progressive", C.c_int32),
        ("restart_interval", C.c_int32), ("mxx", C.c_int32), ("myy", C.c_int32), ("y_stride", C.c_int32),
        ("c_stride", C.c_int32), ("device", C.c_int32), ("reserved", C.c_int32), ("rgba_len", C.c_uint64),
        ("native_len", C.c_uint64), ("native_cb_off", C.c_uint64), ("native_cr_off", C.c_uint64),
    ]


class ZpxParseReport(C.Structure):
    _fields_ = [
        ("status", C.c_int32), ("n_scans", C.c_int32), ("n_intervals", C.c_int32), ("pending_err", C.c_int32),
        ("pending_after_interval", C.c_int32), ("trailing_err", C.c_int32), ("fused", C.c_int32),
        ("mode", C.c_int32), ("entropy_bytes", C.c_uint64), ("stuffed_bytes", C.c_uint64), ("unstuffed_bytes", C.c_uint64),
        ("n_pieces", C.c_int32), ("max_piece", C.c_int32), ("pieces_ok", C.c_int32), ("lane_script", C.c_int32),
    ]


class ZpxTiming(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_float), ("entropy_ms", C.c_float), ("idct_ms", C.c_float), ("total_ms", C.c_float),
        ("d2h_ms", C.c_float), ("entropy_launches", C.c_int32), ("idct_launches", C.c_int32),
        ("entropy_bytes_in", C.c_uint64), ("coef_bytes", C.c_uint64), ("rgba_bytes", C.c_uint64),
        ("pixels", C.c_uint64), ("idct_fused_bytes", C.c_uint64), ("idct_fused_ms", C.c_float),
        ("images", C.c_int32), ("images_failed", C.c_int32),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/zpix_cuda.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("zpx_ctx_create", C.c_int32, [C.POINTER(C.c_int32), C.c_int32, C.POINTER(_P)]),
    ("zpx_ctx_destroy", None, [_P]),
    ("zpx_ctx_num_devices", C.c_int32, [_P]),
    ("zpx_last_cuda_error", C.c_int32, [_P]),
    ("zpx_last_cuda_error_string", C.c_char_p, [_P]),
    ("zpx_batch_open", C.c_int32, [_P, C.POINTER(_P), C.POINTER(C.c_size_t), C.c_int32, C.POINTER(_P)]),
    ("zpx_batch_size", C.c_int32, [_P]),
    ("zpx_batch_info", C.c_int32, [_P, C.c_int32, C.POINTER(ZpxImageInfo)]),
    ("zpx_batch_upload", C.c_int32, [_P]),
    ("zpx_batch_decode", C.c_int32, [_P, _P]),
    ("zpx_batch_fetch_rgba", C.c_int32, [_P, C.POINTER(_P), C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]),
    ("zpx_batch_fetch_native", C.c_int32, [_P, C.POINTER(_P), C.POINTER(C.c_int32)]),
    ("zpx_batch_status", C.c_int32, [_P, C.POINTER(C.c_int32)]),
    ("zpx_batch_device_rgba", _P, [_P, C.c_int32]),
    ("zpx_batch_fetch_coefficients", C.c_int32, [_P, C.c_int32, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    ("zpx_batch_open_synthetic", C.c_int32, [_P, C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int32, C.POINTER(_P)]),
    ("zpx_batch_set_coefficients", C.c_int32, [_P, C.c_int32, _P, C.c_size_t]),
    ("zpx_test_colour", C.c_int32, [_P, C.c_int32, _P, C.c_size_t, _P]),
    ("zpx_batch_timing", C.c_int32, [_P, C.c_int32, C.POINTER(ZpxTiming)]),
    ("zpx_batch_close", None, [_P]),
    ("zpx_decode_batch_rgba", C.c_int32, [_P, C.POINTER(_P), C.POINTER(C.c_size_t), C.c_int32, C.POINTER(_P),
                                          C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]),
    ("zpx_decode_batch_native", C.c_int32, [_P, C.POINTER(_P), C.POINTER(C.c_size_t), C.c_int32, C.POINTER(_P), C.POINTER(C.c_int32)]),
    ("zpx_probe", C.c_int32, [_P, C.c_size_t, C.POINTER(ZpxImageInfo)]),
    ("zpx_parse_report_of", C.c_int32, [_P, C.c_size_t, C.POINTER(ZpxImageInfo), C.POINTER(ZpxParseReport)]),
    ("zpx_partition", C.c_int32, [C.POINTER(C.c_uint64), C.c_int32, C.c_int32, C.POINTER(C.c_int32)]),
    ("zpx_ctx_set_option", C.c_int32, [_P, C.c_int32, C.c_int64]),
    ("zpx_host_alloc", _P, [C.c_size_t]),
    ("zpx_host_free", None, [_P]),
    ("zpx_error_name", C.c_char_p, [C.c_int32]),
    ("zpx_abi_version", C.c_int32, []),
    ("zpx_ctx_kernel_launches", C.c_uint64, [_P]),
]


def _preload_cudart():
    # libzpixcuda.so links libcudart.so.12 dynamically; make sure the copy torch uses is the one loaded
    for base in sys.path:
        for p in glob.glob(os.path.join(base, "nvidia", "cuda_runtime", "lib", "libcudart.so.*")):
            try:
                C.CDLL(p, mode=C.RTLD_GLOBAL)
                return
            except OSError:
                pass


def _load():
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
            "(or python -c 'import __graft_entry__ as g; g.build()'). zpix_b200 has no CPU fallback."
        )
    _preload_cudart()
    lib = C.CDLL(SO_PATH)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()
