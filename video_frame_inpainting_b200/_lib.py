"""ctypes binding of libtai_b200.so (the C ABI declared in include/tai_b200.h).

Plays the role of the reference's cffi loader, src/separable_convolution/_ext/cunnex/__init__.py:1-15
(`_wrap_function` over every exported symbol).  There is no CPU fallback: if the shared library is
missing and cannot be built, importing any operator raises.
"""
from __future__ import annotations

import ctypes
import os
import threading

from . import build as _build

_c_f = ctypes.c_void_p      # device pointer to float (or NULL)
_c_i = ctypes.c_int
_c_s = ctypes.c_void_p      # cudaStream_t

# name -> (restype, argtypes); must list every symbol include/tai_b200.h declares.
SIGNATURES = {
    "tai_b200_abi_version": (_c_i, []),
    "tai_b200_last_error": (ctypes.c_char_p, []),
    "tai_b200_launch_count": (ctypes.c_longlong, []),
    "tai_b200_last_path": (ctypes.c_char_p, []),
    "tai_b200_timing_enable": (_c_i, [_c_i]),
    "tai_b200_timing_report": (_c_i, [ctypes.c_char_p, _c_i]),
    "SeparableConvolution_cuda_forward_b200": (_c_i, [_c_f] * 4 + [_c_i] * 5 + [_c_s]),
    "SeparableConvolution_cuda_backward_b200": (_c_i, [_c_f] * 7 + [_c_i] * 5 + [_c_s]),
    "tai_fused_forward_b200": (_c_i, [_c_f] * 9 + [_c_i] * 5 + [ctypes.c_float] * 2 + [_c_s]),
    "tai_fused_backward_workspace_bytes": (ctypes.c_longlong, [_c_i] * 5),
    "tai_fused_backward_b200": (_c_i, [_c_f] * 16 + [_c_i] * 5 + [ctypes.c_float] * 2 + [_c_s]),
    "replication_pad_forward_b200": (_c_i, [_c_f] * 2 + [_c_i] * 4 + [_c_s]),
    "replication_pad_backward_b200": (_c_i, [_c_f] * 2 + [_c_i] * 4 + [_c_s]),
    "convlstm_gates_forward_b200": (_c_i, [_c_f] * 3 + [_c_i] * 3 + [ctypes.c_float, _c_s]),
    "convlstm_gates_backward_b200": (_c_i, [_c_f] * 5 + [_c_i] * 3 + [ctypes.c_float, _c_s]),
    "flow_warp_forward_b200": (_c_i, [_c_f] * 3 + [_c_i] * 4 + [_c_s]),
    "flow_warp_backward_b200": (_c_i, [_c_f] * 5 + [_c_i] * 4 + [_c_s]),
    "slomo_flow_combine_warp_forward_b200": (_c_i, [_c_f] * 4 + [ctypes.c_double] + [_c_f] * 4 + [_c_i] * 4 + [_c_s]),
    "slomo_refine_blend_forward_b200": (_c_i, [_c_f] * 7 + [ctypes.c_double] + [_c_f] + [_c_i] * 4 + [_c_s]),
    "slomo_interp_input_forward_b200": (_c_i, [_c_f] * 7 + [_c_i] * 5 + [_c_s]),
    "slomo_interp_input_backward_b200": (_c_i, [_c_f] * 9 + [_c_i] * 5 + [_c_s]),
    "slomo_refine_blend_batched_forward_b200": (_c_i, [_c_f] * 8 + [_c_i] * 5 + [_c_s]),
    "slomo_refine_blend_batched_backward_b200": (_c_i, [_c_f] * 13 + [_c_i] * 5 + [_c_s]),
    "upsample_bilinear2x_forward_b200": (_c_i, [_c_f] * 2 + [ctypes.c_longlong] + [_c_i] * 2 + [_c_s]),
    "upsample_bilinear2x_backward_b200": (_c_i, [_c_f] * 2 + [ctypes.c_longlong] + [_c_i] * 2 + [_c_s]),
    "unpool_add_forward_b200": (_c_i, [_c_f] * 3 + [ctypes.c_longlong] + [_c_i] * 2 + [_c_s]),
    "unpool_backward_b200": (_c_i, [_c_f] * 2 + [ctypes.c_longlong] + [_c_i] * 2 + [_c_s]),
    "maxpool2x2_forward_b200": (_c_i, [_c_f] * 3 + [ctypes.c_longlong] + [_c_i] * 2 + [_c_s]),
    "maxpool2x2_backward_b200": (_c_i, [_c_f] * 3 + [ctypes.c_longlong] + [_c_i] * 2 + [_c_s]),
    "l2_gdl_loss_workspace_bytes": (ctypes.c_longlong, [ctypes.c_longlong] + [_c_i] * 2),
    "l2_gdl_loss_forward_b200": (_c_i, [_c_f] * 2 + [ctypes.c_longlong] + [_c_i] * 2 + [ctypes.c_float] * 2 + [_c_f] * 2 + [_c_s]),
    "l2_gdl_loss_backward_b200": (_c_i, [_c_f] * 2 + [ctypes.c_longlong] + [_c_i] * 2 + [ctypes.c_float] * 2 + [_c_f] * 3 + [_c_s]),
    "bias_act_forward_b200": (_c_i, [_c_f] * 2 + [ctypes.c_longlong] + [_c_i] * 3 + [ctypes.c_float, _c_s]),
    "bias_act_backward_workspace_bytes": (ctypes.c_longlong, [ctypes.c_longlong, _c_i]),
    "bias_act_backward_b200": (_c_i, [_c_f] * 5 + [ctypes.c_longlong] + [_c_i] * 3 + [ctypes.c_float, _c_s]),
    "l2_normalize_b200": (_c_i, [_c_f] * 2 + [_c_i, ctypes.c_float, _c_s]),
    "gather_concat_forward_b200": (_c_i, [ctypes.c_void_p, _c_i, _c_f, ctypes.c_longlong, _c_i, _c_i, _c_s]),
    "gather_concat_backward_b200": (_c_i, [ctypes.c_void_p, _c_i, _c_f, ctypes.c_longlong, _c_i, _c_i, _c_s]),
    "gray_difference_frames_b200": (_c_i, [_c_f] * 2 + [_c_i] * 6 + [_c_s]),
    "gray_difference_pair_forward_b200": (_c_i, [_c_f] * 3 + [ctypes.c_longlong] + [_c_i] * 3 + [_c_s]),
    "gray_difference_pair_backward_b200": (_c_i, [_c_f] * 3 + [ctypes.c_longlong] + [_c_i] * 3 + [_c_s]),
    "frames_to_uint8_b200": (_c_i, [_c_f] * 2 + [ctypes.c_longlong] + [_c_i] * 4 + [_c_s]),
    "tai_b200_ffma_probe": (_c_i, [_c_f] + [_c_i] * 4 + [_c_s]),
}

class CatBlock(ctypes.Structure):
    """tai_cat_block of include/tai_b200.h."""
    _fields_ = [("src", ctypes.c_void_p), ("src_sample", ctypes.c_longlong), ("src_sample_stride", ctypes.c_longlong),
                ("dst_sample", ctypes.c_longlong),
                ("dst_sample_stride", ctypes.c_longlong), ("dst_channel", ctypes.c_longlong), ("channels", ctypes.c_int),
                ("samples", ctypes.c_int), ("fill_value", ctypes.c_float)]


ERROR_NAMES = {0: "TAI_OK", -1: "TAI_ERR_INVALID_ARGUMENT", -2: "TAI_ERR_UNSUPPORTED",
               -3: "TAI_ERR_TOO_LARGE", -4: "TAI_ERR_CUDA"}

_lock = threading.Lock()
_handle = None


class TaiB200Error(RuntimeError):
    """A C-ABI entry point returned a non-zero code (the reference surfaced THCudaCheck failures
    as RuntimeError too, kernel.cu:184,212,226,241)."""

    def __init__(self, fn, code, message):
        super().__init__("%s failed: %s (%d): %s" % (fn, ERROR_NAMES.get(code, "?"), code, message))
        self.code = code


def library_path() -> str:
    return _build.LIB_PATH


def load() -> ctypes.CDLL:
    """Load (building first if the sources are newer and nvcc is present) and type every symbol."""
    global _handle
    if _handle is not None:
        return _handle
    with _lock:
        if _handle is not None:
            return _handle
        path = _build.LIB_PATH
        try:
            if not _build.is_up_to_date():
                _build.build_library()
        except Exception as exc:  # no nvcc on this box, or a compile error
            if not os.path.isfile(path):
                raise RuntimeError(
                    "libtai_b200.so is missing and could not be built (%s). The TAI hot path has no "
                    "CPU or PyTorch fallback; run `python -m video_frame_inpainting_b200.build`." % exc)
        lib = ctypes.CDLL(path)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header and library disagree
            fn.restype = restype
            fn.argtypes = argtypes
        got = lib.tai_b200_abi_version()
        if got != 1:
            raise RuntimeError("libtai_b200.so ABI version %d, expected 1" % got)
        _handle = lib
    return _handle


def call(name: str, *args) -> None:
    """Invoke an int-returning entry point and raise TaiB200Error on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise TaiB200Error(name, rc, lib.tai_b200_last_error().decode("utf-8", "replace"))


def launch_count() -> int:
    return int(load().tai_b200_launch_count())


def last_path() -> str:
    """Kernel family of this thread's last separable-convolution launch ("fwd:v3", "bwd_i:v4", "fwd:tiled", ...)."""
    return (load().tai_b200_last_path() or b"").decode()


def timing_enable(on: bool) -> None:
    """Bracket every kernel launch of the library with CUDA events (bench.py's roofline measurement)."""
    call("tai_b200_timing_enable", 1 if on else 0)


def timing_report():
    """[{name, launches, ms, flops, bytes}] for the launches recorded since timing_enable(True)."""
    import json
    buf = ctypes.create_string_buffer(1 << 16)
    call("tai_b200_timing_report", buf, len(buf))
    return json.loads(buf.value.decode())
