"""ctypes binding of the C-ABI library (include/mixdq_b200.h).

There is no CPU fallback: if the library cannot be loaded the ops raise. Loading never
triggers a build implicitly on a machine without nvcc — `__graft_entry__.build()` (or
`python -m mixdq_b200.build`) produces the .so in-tree.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_int, c_int32, c_int64, c_void_p, POINTER
from pathlib import Path

_LIB = None
ABI_VERSION = 2
LIB_PATH = Path(__file__).resolve().parent / "libmixdq_b200.so"

# every symbol include/mixdq_b200.h declares (checked by tests/test_abi.py)
ABI_SYMBOLS = [
    "mixdq_abi_version", "mixdq_set_workspace", "mixdq_stream_capture_id", "mixdq_debug_set_pdl", "mixdq_strerror", "mixdq_last_path", "mixdq_force_simt",
    "mixdq_debug_force_bn", "mixdq_debug_force_splits", "mixdq_debug_set_persist", "mixdq_debug_set_persist_bn", "mixdq_debug_set_conv_halo", "mixdq_debug_set_timing_buffer",
    "mixdq_debug_set_mode", "mixdq_debug_set_cluster", "mixdq_debug_set_two_pass", "mixdq_debug_set_quant_timing_buffer",
    "mixdq_quant_i8_static", "mixdq_quant_i8_static_strided", "mixdq_quant_i8_nchw2nhwc",
    "mixdq_quant_dynamic_ws_bytes", "mixdq_quant_i8_dynamic",
    "mixdq_gemm_w8a8_f16", "mixdq_gemm_w8a8_f16_dyn", "mixdq_gemm_w4a8_f16",
    "mixdq_conv_w8a8_f16", "mixdq_conv1x1_split_w8a8_f16",
    "mixdq_gemm_w8a8_f16_dyn_res", "mixdq_conv_w8a8_f16_dyn", "mixdq_conv1x1_split_w8a8_f16_dyn",
    "mixdq_quant_i8_dynamic_rows", "mixdq_ln_quant_i8_dynamic", "mixdq_geglu_quant_i8_dynamic", "mixdq_gn_quant_i8_dynamic",
    "mixdq_gemm_w8a8_geglu_f16_dyn", "mixdq_quant_i8_premm",
    "mixdq_gemm_w4a8_f16_dyn_res", "mixdq_gemm_w4a8_geglu_f16_dyn", "mixdq_conv_w4a8_f16",
    "mixdq_conv_w4a8_f16_dyn", "mixdq_quant_i8_dynamic_bits", "mixdq_quant_i8_static_range", "mixdq_minmax_f16",
    "mixdq_ln_quant_i8_static", "mixdq_gn_quant_i8_static", "mixdq_gemm_geglu_i8_static",
]


class MixdqLibraryError(RuntimeError):
    pass


def _declare(lib: ctypes.CDLL) -> None:
    P = c_void_p
    lib.mixdq_abi_version.restype = c_int
    lib.mixdq_abi_version.argtypes = []
    lib.mixdq_set_workspace.restype = c_int
    lib.mixdq_set_workspace.argtypes = [c_int, P, P, c_int64]
    lib.mixdq_stream_capture_id.restype = c_int
    lib.mixdq_stream_capture_id.argtypes = [P, POINTER(ctypes.c_uint64)]
    lib.mixdq_debug_set_pdl.restype = None
    lib.mixdq_debug_set_pdl.argtypes = [c_int]
    lib.mixdq_strerror.restype = c_char_p
    lib.mixdq_strerror.argtypes = [c_int]
    lib.mixdq_last_path.restype = c_char_p
    lib.mixdq_last_path.argtypes = []
    lib.mixdq_force_simt.restype = None
    lib.mixdq_force_simt.argtypes = [c_int]
    lib.mixdq_debug_force_bn.restype = None
    lib.mixdq_debug_force_bn.argtypes = [c_int]
    lib.mixdq_debug_set_persist.restype = None
    lib.mixdq_debug_set_persist.argtypes = [c_int, c_int]
    lib.mixdq_debug_set_persist_bn.restype = None
    lib.mixdq_debug_set_persist_bn.argtypes = [c_int]
    lib.mixdq_debug_set_conv_halo.restype = None
    lib.mixdq_debug_set_conv_halo.argtypes = [c_int]
    lib.mixdq_quant_i8_dynamic_bits.restype = c_int
    lib.mixdq_quant_i8_dynamic_bits.argtypes = [P, c_int64, c_int64, c_int64, c_int, P, P, P, P, P]
    lib.mixdq_minmax_f16.restype = c_int
    lib.mixdq_minmax_f16.argtypes = [P, c_int64, P, P, P]
    lib.mixdq_quant_i8_static_range.restype = c_int
    lib.mixdq_quant_i8_static_range.argtypes = [P, c_int64, P, P, c_int, c_int, P, P]
    lib.mixdq_debug_force_splits.restype = None
    lib.mixdq_debug_force_splits.argtypes = [c_int]
    lib.mixdq_debug_set_timing_buffer.restype = None
    lib.mixdq_debug_set_timing_buffer.argtypes = [P]
    lib.mixdq_debug_set_quant_timing_buffer.restype = c_int
    lib.mixdq_debug_set_quant_timing_buffer.argtypes = [P, P, P]
    lib.mixdq_debug_set_two_pass.restype = None
    lib.mixdq_debug_set_two_pass.argtypes = [c_int]
    lib.mixdq_debug_set_cluster.restype = None
    lib.mixdq_debug_set_cluster.argtypes = [c_int]
    lib.mixdq_debug_set_mode.restype = None
    lib.mixdq_debug_set_mode.argtypes = [c_int]

    lib.mixdq_quant_i8_static.restype = c_int
    lib.mixdq_quant_i8_static.argtypes = [P, c_int64, P, P, P, P]
    lib.mixdq_quant_i8_static_strided.restype = c_int
    lib.mixdq_quant_i8_static_strided.argtypes = [P, c_int64, c_int64, c_int64, c_int64, c_int64,
                                                  P, P, P, c_int64, P]
    lib.mixdq_quant_i8_nchw2nhwc.restype = c_int
    lib.mixdq_quant_i8_nchw2nhwc.argtypes = [P, c_int, c_int, c_int, c_int, POINTER(c_int64),
                                             c_int, c_int, P, P, P, P]
    lib.mixdq_quant_dynamic_ws_bytes.restype = c_int64
    lib.mixdq_quant_dynamic_ws_bytes.argtypes = []
    lib.mixdq_quant_i8_dynamic.restype = c_int
    lib.mixdq_quant_i8_dynamic.argtypes = [P, c_int64, P, P, P, P, P]

    lib.mixdq_gemm_w8a8_f16.restype = c_int
    lib.mixdq_gemm_w8a8_f16.argtypes = [P, c_int64, P, P, P, P, P, c_int64, c_int, c_int, c_int,
                                        P, P]
    lib.mixdq_gemm_w8a8_f16_dyn.restype = c_int
    lib.mixdq_gemm_w8a8_f16_dyn.argtypes = [P, c_int64, P, P, P, P, P, P, P, c_int64, c_int,
                                            c_int, c_int, P, P]
    lib.mixdq_gemm_w4a8_f16.restype = c_int
    lib.mixdq_gemm_w4a8_f16.argtypes = [P, c_int64, P, P, P, P, P, c_int64, c_int, c_int, c_int,
                                        P, P]
    lib.mixdq_conv_w8a8_f16.restype = c_int
    lib.mixdq_conv_w8a8_f16.argtypes = [P, c_int64, P, P, P, P, P, P, P,
                                        c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                        c_int, P, P]
    lib.mixdq_conv1x1_split_w8a8_f16.restype = c_int
    lib.mixdq_conv1x1_split_w8a8_f16.argtypes = [P, c_int64, P, c_int, P, P,
                                                 P, c_int64, P, c_int, P, P,
                                                 P, P, c_int64, c_int, c_int, P]


    from ctypes import c_float
    lib.mixdq_gemm_w8a8_f16_dyn_res.restype = c_int
    lib.mixdq_gemm_w8a8_f16_dyn_res.argtypes = [P, c_int64, P, P, P, P, P, P, P, c_int64, P,
                                                c_int64, c_int, c_int, c_int, P, P]
    lib.mixdq_conv_w8a8_f16_dyn.restype = c_int
    lib.mixdq_conv_w8a8_f16_dyn.argtypes = [P, c_int64, P, P, P, P, P, P, P, P, c_int64, P, P,
                                            c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                            c_int, c_int, P, P]
    lib.mixdq_conv1x1_split_w8a8_f16_dyn.restype = c_int
    lib.mixdq_conv1x1_split_w8a8_f16_dyn.argtypes = [P, c_int64, P, c_int, P, P, P, P,
                                                     P, c_int64, P, c_int, P, P, P, P,
                                                     P, P, c_int64, P, c_int64, c_int, c_int, P]
    lib.mixdq_quant_i8_dynamic_rows.restype = c_int
    lib.mixdq_quant_i8_dynamic_rows.argtypes = [P, c_int64, c_int, c_int, P, P, P, P, P]
    lib.mixdq_ln_quant_i8_dynamic.restype = c_int
    lib.mixdq_ln_quant_i8_dynamic.argtypes = [P, c_int64, c_int, c_int, P, P, c_float, P, P, P, P,
                                              P, P]
    lib.mixdq_geglu_quant_i8_dynamic.restype = c_int
    lib.mixdq_geglu_quant_i8_dynamic.argtypes = [P, c_int64, c_int, c_int, P, P, P, P, P, P]
    lib.mixdq_gemm_w8a8_geglu_f16_dyn.restype = c_int
    lib.mixdq_gemm_w8a8_geglu_f16_dyn.argtypes = [P, c_int64, P, P, P, P, P, P, P, c_int64,
                                                  c_int, c_int, c_int, P, P]
    for w8, w4 in (("mixdq_gemm_w8a8_f16_dyn_res", "mixdq_gemm_w4a8_f16_dyn_res"),
                   ("mixdq_gemm_w8a8_geglu_f16_dyn", "mixdq_gemm_w4a8_geglu_f16_dyn"),
                   ("mixdq_conv_w8a8_f16", "mixdq_conv_w4a8_f16"),
                   ("mixdq_conv_w8a8_f16_dyn", "mixdq_conv_w4a8_f16_dyn")):
        getattr(lib, w4).restype = c_int              # same signatures, packed weight pointer
        getattr(lib, w4).argtypes = getattr(lib, w8).argtypes
    lib.mixdq_quant_i8_premm.restype = c_int
    lib.mixdq_quant_i8_premm.argtypes = [P, c_int64, P, P, P, P, P]
    lib.mixdq_gn_quant_i8_dynamic.restype = c_int
    lib.mixdq_gn_quant_i8_dynamic.argtypes = [P, c_int64, c_int, c_int, c_int, c_int, P, P,
                                              c_float, c_int, P, P, P, P, P, P]
    lib.mixdq_ln_quant_i8_static.restype = c_int
    lib.mixdq_ln_quant_i8_static.argtypes = [P, c_int64, c_int, c_int, P, P, c_float, P, P, P, P, P]
    lib.mixdq_gn_quant_i8_static.restype = c_int
    lib.mixdq_gn_quant_i8_static.argtypes = [P, c_int64, c_int, c_int, c_int, c_int, P, P,
                                             c_float, c_int, P, P, P, P, P]
    lib.mixdq_gemm_geglu_i8_static.restype = c_int
    lib.mixdq_gemm_geglu_i8_static.argtypes = [P, c_int64, P, c_int, P, P, P, P, P, P, P, P,
                                               c_int64, c_int, c_int, c_int, P, P]


def load() -> ctypes.CDLL:
    """Load libmixdq_b200.so (once). Raises MixdqLibraryError if it is missing."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = Path(os.environ.get("MIXDQ_B200_LIB", LIB_PATH))
    if not path.exists():
        raise MixdqLibraryError(
            f"{path} not found: the CUDA extension is not built. Run "
            "`python -m mixdq_b200.build` (needs nvcc). There is no CPU fallback.")
    try:
        lib = ctypes.CDLL(str(path))
    except OSError as e:  # pragma: no cover - depends on the machine
        raise MixdqLibraryError(f"cannot load {path}: {e}") from e
    for sym in ABI_SYMBOLS:
        if not hasattr(lib, sym):
            raise MixdqLibraryError(f"{path} does not export {sym}")
    _declare(lib)
    if lib.mixdq_abi_version() != ABI_VERSION:
        raise MixdqLibraryError("ABI version mismatch")
    _LIB = lib
    return lib


def check(code: int) -> None:
    """Turn a non-zero status into the RuntimeError the reference's TORCH_CHECK would raise."""
    if code != 0:
        msg = load().mixdq_strerror(code).decode()
        raise RuntimeError(msg)


def capture_id(stream: int) -> int:
    """0 outside CUDA-graph capture, else the id of the capture sequence `stream` belongs to."""
    out = ctypes.c_uint64(0)
    check(load().mixdq_stream_capture_id(stream, ctypes.byref(out)))
    return int(out.value)


def last_path() -> str:
    return load().mixdq_last_path().decode()
