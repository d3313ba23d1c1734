from .SeparableConvolution import SeparableConvolution  # noqa: F401
