"""B200-native (sm_100a) implementation of the TAI / bi-TAI video-frame-inpainting hot path.

Python host code over a C-ABI CUDA library (include/tai_b200.h -> lib/libtai_b200.so).  Mirrors the
reference's operator and model interfaces for that path (MichiganCOG/video-frame-inpainting:
src/separable_convolution, src/models/{tai,mcnet,slomo}); see DESIGN.md.
"""
__all__ = ["ops", "build"]
