from .conv2d import QuantizedConv2d  # noqa: F401
from .linear import QuantizedLinear  # noqa: F401
