"""Representative launches of every kernel family of the batch-1 SDXL step, for one
`ncu --set full` capture (each family twice: the second launch of a family is the warm one)."""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
ops.DYNAMIC_QUANT_CACHE = False
g = torch.Generator().manual_seed(0)


def weights(N, K):
    w = torch.randint(-127, 128, (N, K), dtype=torch.int8, generator=g).to(dev)
    return w, (0.001 + 0.01 * torch.rand(N, generator=g)).to(dev), w.float().sum(1)


M, C = 256, 1280
x = torch.randn(M, C, generator=g).half().to(dev)
gamma = torch.ones(C, device=dev).half(); beta = torch.zeros(C, device=dev).half()
w_o, s_o, sum_o = weights(C, C)              # to_out / to_q      (BN=32, dual issue)
w_qkv, s_qkv, sum_qkv = weights(3 * C, C)    # q/k/v concatenated (BN=64)
w_ff, s_ff, sum_ff = weights(8 * C, C)       # ff.net.0.proj      (GEGLU epilogue, BN=160)
w_f2, s_f2, sum_f2 = weights(C, 4 * C)       # ff.net.2           (split-K)
idx = ops.geglu_interleave_index(4 * C, dev)
w_ffi, s_ffi, sum_ffi = w_ff[idx].contiguous(), s_ff[idx].contiguous(), sum_ff[idx].contiguous()
img = torch.randn(1, 320, 64, 64, generator=g).half().to(dev).contiguous(memory_format=torch.channels_last)
gw = torch.ones(320, device=dev).half(); gb = torch.zeros(320, device=dev).half()
wc = torch.randint(-127, 128, (320, 320, 3, 3), dtype=torch.int8, generator=g).to(dev).contiguous(
    memory_format=torch.channels_last)
sc = (0.001 + 0.01 * torch.rand(320, generator=g)).to(dev)
wsum_c = wc.float().sum(1, keepdim=True).contiguous()

for rep in range(2):
    q8, s, z = ops.layernorm_quantize_dynamic(x, gamma, beta, 1e-5)
    qkv = ops.qlinear_dynamic_fused(q8, w_qkv, s_qkv, s, z, sum_qkv, None)
    o8, s2, z2 = ops.quantize_per_tensor_dynamic(qkv[:, :C].contiguous())
    y = ops.qlinear_dynamic_fused(o8, w_o, s_o, s2, z2, sum_o, None, residual=x)
    q8, s, z = ops.layernorm_quantize_dynamic(y, gamma, beta, 1e-5)
    g8, s3, z3 = ops.qlinear_geglu_quantize_dynamic(q8, w_ffi, s_ffi, s, z, sum_ffi, None)
    y2 = ops.qlinear_dynamic_fused(g8, w_f2, s_f2, s3, z3, sum_f2, None, residual=y)
    h8, s4, z4 = ops.groupnorm_quantize_dynamic(img, 32, gw, gb, 1e-5, True)
    c = ops.qconv2d_dynamic_fused(h8, wc, sc, s4, z4, wsum_c, None, None, 1, 1, None, img)
torch.cuda.synchronize()
print("ok", float(y2.float().abs().mean()), float(c.float().abs().mean()))
