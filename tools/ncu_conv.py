"""A few launches of one 3x3 convolution on the persistent kernel (halo form unless
MIXDQ_CONV_HALO=0), for one `ncu --set full` capture: python tools/ncu_conv.py n,hw,c,k"""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
n, h, c, k = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "64,16,1280,1280").split(","))
x = torch.randint(-128, 128, (n, c, h, h), dtype=torch.int8, device=dev).contiguous(memory_format=torch.channels_last)
ws = [torch.randint(-127, 128, (k, c, 3, 3), dtype=torch.int8, device=dev).contiguous(memory_format=torch.channels_last)
      for _ in range(3)]
sc = torch.ones(k, device=dev); s1 = torch.tensor(1.0, device=dev); zp = torch.tensor(3.0, device=dev)
for w in ws:
    y = ops.qconv2d_w8_a8_ohalf(x, w, sc, s1, zp, sc, w.float().sum(1, keepdim=True).contiguous(), None, None, 1, 1, 1)
torch.cuda.synchronize()
print("ok", lib.mixdq_last_path().decode(), float(y.float().abs().mean()))
