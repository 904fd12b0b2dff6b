"""Cross-attention segment of a transformer block as a dependent chain inside a CUDA graph:
to_q GEMM -> attention -> quantise -> to_out GEMM(+residual), own kernel vs library SDPA."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
ops.DYNAMIC_QUANT_CACHE = False
g = torch.Generator().manual_seed(0)
T, C, H, Lk = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), 77) if len(sys.argv) > 3 else (256, 1280, 20, 77)
w = torch.randint(-127, 128, (C, C), dtype=torch.int8, generator=g).to(dev)
wsc = (0.001 + 0.01 * torch.rand(C, generator=g)).to(dev); wsum = w.float().sum(1)
x = torch.randn(1, T, C, generator=g).half().to(dev)
kv = torch.randn(1, Lk, 4 * C, generator=g).half().to(dev)
k, v = kv[..., :C], kv[..., C:2 * C]
q8, s, z = ops.quantize_per_tensor_dynamic(x)
NB = 10


def heads(t):
    b, n, c = t.shape
    return t.view(b, n, H, 64).transpose(1, 2)


def chain(custom):
    y = x
    for _ in range(NB):
        q = ops.qlinear_dynamic_fused(q8, w, wsc, s, z, wsum, None)
        if custom:
            o8, s2, z2 = ops.cross_attention_quantize_dynamic(q, k, v, H)
        else:
            o = F.scaled_dot_product_attention(heads(q), heads(k), heads(v)).transpose(1, 2).reshape(1, T, C)
            o8, s2, z2 = ops.quantize_per_tensor_dynamic(o)
        y = ops.qlinear_dynamic_fused(o8, w, wsc, s2, z2, wsum, None, residual=y)
    return y


for custom in (True, False, True, False):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        chain(custom)
    torch.cuda.current_stream().wait_stream(side)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = chain(custom)
    for _ in range(3):
        gr.replay()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    print(f"T={T} C={C} custom={custom}: {e0.elapsed_time(e1) / 20 / NB * 1e3:.2f} us per segment", flush=True)
