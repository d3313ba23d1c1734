"""ctypes harness for oracle/_ref/libsepconv_ref.so -- the reference's OWN native code
(SeparableConvolution_kernel.cu + SeparableConvolution_cuda.c, unmodified, compiled for sm_100a by
oracle/Makefile against the THC stand-in headers in oracle/ref_shim/).  Test infrastructure only."""
import ctypes
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libsepconv_ref.so")


class THCudaTensor(ctypes.Structure):  # mirrors oracle/ref_shim/THC.h
    _fields_ = [("data", ctypes.c_void_p), ("size", ctypes.c_long * 4), ("stride", ctypes.c_long * 4)]


def available():
    return os.path.isfile(REF_LIB)


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(REF_LIB)
        P = ctypes.POINTER(THCudaTensor)
        _lib.SeparableConvolution_cuda_forward.argtypes = [P] * 4 + [ctypes.c_int]
        _lib.SeparableConvolution_cuda_forward.restype = ctypes.c_int
        _lib.SeparableConvolution_cuda_backward.argtypes = [P] * 7 + [ctypes.c_int]
        _lib.SeparableConvolution_cuda_backward.restype = ctypes.c_int
        _lib.ref_shim_set_stream.argtypes = [ctypes.c_void_p]
        _lib.ref_shim_last_error.restype = ctypes.c_int
    return _lib


def _wrap(t):
    assert t.is_cuda and t.is_contiguous() and t.dim() == 4
    s = THCudaTensor()
    s.data = t.data_ptr()
    for k in range(4):
        s.size[k] = t.size(k)
        s.stride[k] = t.stride(k)
    return s


def _sync_stream():
    import torch
    lib().ref_shim_set_stream(torch.cuda.current_stream().cuda_stream)


def forward(input, vertical, horizontal, ks):
    """The reference call sequence of SeparableConvolution.py:36-46 (zero-filled output, then FFI)."""
    import torch
    B, C = input.shape[:2]
    out = torch.zeros(B, C, vertical.size(2), vertical.size(3), device=input.device)
    _sync_stream()
    args = [_wrap(t) for t in (input, vertical, horizontal, out)]
    rc = lib().SeparableConvolution_cuda_forward(*[ctypes.byref(a) for a in args], ks)
    assert rc == 1 and lib().ref_shim_last_error() == 0  # the reference always returns 1 (cuda.c:24)
    return out


def backward(grad_output, input, vertical, horizontal, ks):
    """SeparableConvolution.py:69-84."""
    import torch
    gi, gv, gh = torch.zeros_like(input), torch.zeros_like(vertical), torch.zeros_like(horizontal)
    _sync_stream()
    args = [_wrap(t) for t in (grad_output, input, vertical, horizontal, gi, gv, gh)]
    rc = lib().SeparableConvolution_cuda_backward(*[ctypes.byref(a) for a in args], ks)
    assert rc == 1 and lib().ref_shim_last_error() == 0
    return gi, gv, gh
