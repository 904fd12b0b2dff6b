"""mixdq_b200 — B200-native (sm_100a) implementation of MixDQ's quantized-UNet hot path.

Layout:
  csrc/            CUDA kernels + the C ABI (include/mixdq_b200.h)  -> libmixdq_b200.so
  ops.py           host side of the reference's `mixdq_extension._C` op layer
  nn/              QuantizedLinear / QuantizedConv2d (reference kernels/mixdq_extension/nn)
  quantize.py      convert / swap_module             (reference kernels/quantize.py)
  mixdq.py         quantize_unet, cuda_graph_opt, ComfyUI nodes (reference kernels/mixdq.py)
  unet.py          SDXL-Turbo / SD-Turbo UNet skeletons (diffusers is not available offline)
  dp.py            batch-sharded data-parallel driver (new; the reference is single-GPU)
"""
__version__ = "0.1.0"
