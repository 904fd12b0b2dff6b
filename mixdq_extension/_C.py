"""`mixdq_extension._C`: the five ops of the reference's pybind module
(reference kernels/mixdq_extension/csrc/main.cpp:9-13, quantize.cc:55-62, qlinear.cc:208-234,
qconv2d.cc:208-235), same names and argument order, backed by the C ABI in include/mixdq_b200.h."""
from mixdq_b200.ops import (  # noqa: F401
    qconv2d_w8_a8_ohalf,
    qlinear_fp_reference,
    qlinear_w8_a8_ohalf,
    quantize_per_tensor_to_int8,
    quantize_per_tensor_to_int8_vectorized,
)
