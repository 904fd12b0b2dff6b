"""A few launches of the persistent tcgen05 GEMM at tensor-bound shapes, for one `ncu --set full`
capture (see profiles/README.md): python tools/ncu_persist.py M,N,K [bn] [cs]"""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
M, N, K = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "8192,5120,640").split(","))
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
cs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lib.mixdq_debug_set_persist(2, cs); lib.mixdq_debug_set_persist_bn(bn)
a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(3)]
z = torch.zeros(N, device=dev); o = torch.ones(N, device=dev); s1 = torch.tensor(1.0, device=dev)
for w in ws:
    y = ops.qlinear_w8_a8_ohalf(a, w, o, s1, s1, z, o, z, None)
torch.cuda.synchronize()
print("ok", lib.mixdq_last_path().decode(), float(y.float().abs().mean()))
