"""A few launches of the memory-bound producers at big-batch shapes (SD-Turbo batch 64, SDXL batch
8), for one `ncu --set full` capture:  python tools/ncu_elem.py [sd64|b8|b1]"""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
which = sys.argv[1] if len(sys.argv) > 1 else "sd64"
N, H, C, T = {"sd64": (64, 64, 320, 4096), "b8": (8, 32, 640, 1024), "b1": (1, 16, 1280, 256)}[which]
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(N, C, H, H, generator=g, device=dev, dtype=torch.float16).contiguous(
    memory_format=torch.channels_last)
w = torch.ones(C, device=dev, dtype=torch.float16); b = torch.zeros(C, device=dev, dtype=torch.float16)
inv = torch.tensor(30.0, device=dev); zp = torch.tensor(-3.0, device=dev)
tok = torch.randn(N * T, C, generator=g, device=dev, dtype=torch.float16)
for _ in range(2):
    q1 = ops.groupnorm_quantize_static(x, 32, w, b, 1e-5, True, inv, zp)       # stats + apply
    q2, s, z = ops.groupnorm_quantize_dynamic(x, 32, w, b, 1e-5, True)         # stats + apply + quantise
    q3 = ops.layernorm_quantize_static(tok, w, b, 1e-5, inv, zp)
    q4, s, z = ops.layernorm_quantize_dynamic(tok, w, b, 1e-5)
    q5 = ops.quantize_per_tensor_to_int8(tok, inv, zp)
    q6, s, z = ops.quantize_per_tensor_dynamic(tok.clone())
torch.cuda.synchronize()
print("ok", which, int(q1.float().abs().sum()), int(q3.float().abs().sum()))
