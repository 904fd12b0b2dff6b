"""reference kernels/mixdq_extension/op/qlinear.py:5-6"""
import mixdq_extension._C
from .quant import quantize_per_tensor, quantize_per_tensor_vectorized  # noqa: F401

qlinear = mixdq_extension._C.qlinear_w8_a8_ohalf
qlinear_ref = mixdq_extension._C.qlinear_fp_reference
