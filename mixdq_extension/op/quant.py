"""reference kernels/mixdq_extension/op/quant.py:4-5"""
import mixdq_extension._C

quantize_per_tensor = mixdq_extension._C.quantize_per_tensor_to_int8
quantize_per_tensor_vectorized = mixdq_extension._C.quantize_per_tensor_to_int8_vectorized
