"""Dynamic quantisation (min/max pass + quantise pass) on the tensor sizes of the batch-1 / batch-8
step: a few launches per size, for `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,...`."""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops

dev = torch.device("cuda:0")
_lib.load()
ops.DYNAMIC_QUANT_CACHE = False
sizes = [(256, 1280), (256, 5120), (2048, 1280), (2048, 5120), (8192, 640), (8192, 2560), (32768, 320)]
for (m, c) in sizes:
    xs = [torch.randn(m, c, device=dev, dtype=torch.float16) for _ in range(3)]
    for x in xs * 2:
        q, s, z = ops.quantize_per_tensor_dynamic(x)
torch.cuda.synchronize()
print("ok")
