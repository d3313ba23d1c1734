"""Drop-in for the reference operator ``src/separable_convolution/SeparableConvolution.py``.

Same name, same call (``SeparableConvolution.apply(input, vertical, horizontal, ks)``, bound at
tai.py:172 / twi.py:174), same shape asserts (SeparableConvolution.py:27-33), same return arity of
backward ``(grad_input, grad_vertical, grad_horizontal, None)`` (:89), CPU tensors raise
``NotImplementedError`` (:48-49,86-87).  What changed underneath: the cffi extension
``_ext.cunnex`` is replaced by the sm_100a library ``libtai_b200.so`` (include/tai_b200.h), outputs
are no longer zero-filled before the call (the kernels overwrite every element) and the three
backward kernels of the reference are two (gV+gH fused, gI).
"""
from ..ops import SeparableConvolutionFunction


class SeparableConvolution(SeparableConvolutionFunction):
    """``SeparableConvolution.apply(input, vertical, horizontal, ks=51) -> output``

    input [B,C,Hi,Wi], vertical/horizontal [B,ks,Ho,Wo] with Hi-ks == Ho-1, Wi-ks == Wo-1;
    output[b,c,y,x] = sum_i sum_j input[b,c,y+i,x+j] * vertical[b,i,y,x] * horizontal[b,j,y,x].
    """
