"""CPU-side checks of the boundary: the shared library loads and exports every symbol that
include/tai_b200.h declares (no compute calls -- there is no GPU here), the Python binding lists the same
set, and the operator keeps the reference's CPU behaviour (NotImplementedError, shape asserts)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "tai_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+char\s*\*|long long|int)\s*(\w+)\s*\(", text, flags=re.M)
    assert len(names) >= 18, names
    return set(names)


def test_library_exports_every_declared_symbol():
    from video_frame_inpainting_b200 import _lib, build
    path = build.build_library()
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    for name in declared:
        assert getattr(lib, name) is not None
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib.tai_b200_abi_version.restype = ctypes.c_int
    assert lib.tai_b200_abi_version() == 1


def test_argument_validation_returns_codes_without_a_gpu():
    """Validation happens before any CUDA call, so it can be exercised here: bad arguments come back as
    negative codes with a message, nothing is thrown across the C boundary."""
    from video_frame_inpainting_b200 import _lib
    lib = _lib.load()
    rc = lib.SeparableConvolution_cuda_forward_b200(None, None, None, None, 1, 1, 60, 60, 51, None)
    assert rc == -1 and b"null pointer" in lib.tai_b200_last_error()
    one = ctypes.c_void_p(16)
    rc = lib.SeparableConvolution_cuda_forward_b200(one, one, one, one, 1, 1, 10, 60, 51, None)  # Hi < ks
    assert rc == -1
    rc = lib.tai_fused_forward_b200(one, one, one, one, one, one, one, None, None, 1, 1, 8, 8, 4, 0.5, 0.5, None)
    assert rc == -1 and b"odd" in lib.tai_b200_last_error()
    rc = lib.SeparableConvolution_cuda_forward_b200(one, one, one, one, 4096, 64, 4096, 4096, 51, None)
    assert rc == -3  # >= 2^31 elements
    assert lib.tai_fused_backward_workspace_bytes(2, 3, 8, 8, 5) == 4 * (2 * 2 * 3 * 64 + 2 * 3 * 12 * 12)


def test_operator_keeps_reference_cpu_behaviour():
    from video_frame_inpainting_b200.separable_convolution import SeparableConvolution
    ks = 5
    inp, v, h = torch.randn(1, 1, 12, 12), torch.randn(1, ks, 8, 8), torch.randn(1, ks, 8, 8)
    with pytest.raises(NotImplementedError):  # SeparableConvolution.py:48-49
        SeparableConvolution.apply(inp, v, h, ks)
    from video_frame_inpainting_b200 import ops
    with pytest.raises(AssertionError):       # SeparableConvolution.py:27
        ops.sepconv_shapes(torch.randn(1, 1, 11, 12), v, h, ks)
    with pytest.raises(AssertionError):       # SeparableConvolution.py:29
        ops.sepconv_shapes(inp, v, h, 3)
    with pytest.raises(NotImplementedError):
        ops.convlstm_gates_forward(torch.randn(1, 8, 2, 2), torch.randn(1, 4, 2, 2))
    with pytest.raises(NotImplementedError):
        ops.flow_warp_forward(torch.randn(1, 3, 4, 4), torch.randn(1, 2, 4, 4))


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under video_frame_inpainting_b200/ may import, load or execute
    anything under oracle/ (a product path through the oracle would void every parity claim)."""
    pkg = os.path.join(ROOT, "video_frame_inpainting_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|liboracle|oracle/_ref|oracle\._", text, flags=re.M):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_new_entry_points_validate_arguments_without_a_gpu():
    from video_frame_inpainting_b200 import _lib
    lib = _lib.load()
    one = ctypes.c_void_p(16)
    assert lib.maxpool2x2_forward_b200(one, one, one, 1, 1, 8, None) == -1          # H < 2
    assert lib.maxpool2x2_backward_b200(one, None, one, 1, 4, 4, None) == -1        # no code buffer
    assert lib.l2_gdl_loss_forward_b200(one, one, 1, 4, 4, 1.0, 0.5, one, None, None) == -1   # no workspace
    assert lib.l2_gdl_loss_workspace_bytes(160, 128, 128) > 0
    assert lib.bias_act_forward_b200(one, one, 1, 1, 4, 7, 0.0, None) == -1         # unknown activation
    assert lib.bias_act_backward_b200(one, None, one, one, one, 1, 1, 4, 1, 0.0, None) == -1  # relu needs the output
    assert lib.bias_act_backward_workspace_bytes(64, 64) >= 4 * 64
    assert lib.upsample_bilinear2x_backward_b200(one, one, 1 << 40, 64, 64, None) == -3
    assert lib.frames_to_uint8_b200(one, one, 0, 3, 4, 4, 1, None) == -1
    assert lib.l2_normalize_b200(one, one, 0, 1e-12, None) == -1
