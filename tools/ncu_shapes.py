"""A handful of representative launches for `ncu` (GEMM / conv shapes of the B=1 SDXL step)."""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()


def lin(M, N, K, bn=0, s=0, reps=3):
    lib.mixdq_debug_force_bn(bn); lib.mixdq_debug_force_splits(s)
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(reps)]
    z = torch.zeros(N, device=dev); o = torch.ones(N, device=dev); s1 = torch.tensor(1.0, device=dev)
    for w in ws:
        ops.qlinear_w8_a8_ohalf(a, w, o, s1, s1, z, o, z, None)
    torch.cuda.synchronize()


lin(256, 10240, 1280, 256, 1)
lin(256, 1280, 1280, 64, 1)
lin(256, 1280, 1280, 128, 4)
lin(2048, 10240, 1280, 256, 1)
x = torch.randn(4096 * 320, device=dev).half()
for _ in range(3):
    ops.quantize_per_tensor_dynamic(x)
torch.cuda.synchronize()
