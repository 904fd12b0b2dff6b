"""The hoisted cross-attention K/V projection of SDXL (all 140 to_k / to_v layers as ONE GEMM:
M = 77 context tokens, N = 166400, K = 2048, 341 MB of int8 weights): one-tile kernel vs the
persistent kernel (single CTAs / pairs need two m-tiles), back to back in a graph."""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops
from tools.tops_sweep import graph_time
dev = torch.device("cuda:0")
lib = _lib.load()
M, N, K = 77, 166400, 2048
a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(2)]
z = torch.zeros(N, device=dev); o = torch.ones(N, device=dev); s1 = torch.tensor(1.0, device=dev)
for tag, mode, cs, bn in (("tile", 0, 2, 0), ("persist cs1 bn256", 2, 1, 256), ("persist cs1 bn160", 2, 1, 160),
                          ("persist cs1 bn128", 2, 1, 128), ("heuristic", 1, 2, 0)):
    lib.mixdq_debug_set_persist(mode, cs); lib.mixdq_debug_set_persist_bn(bn)
    keep = []
    t = graph_time([(lambda w=w: keep.append(ops.qlinear_w8_a8_ohalf(a, w, o, s1, s1, z, o, z, None))) for w in ws] * 2)
    print(f"{tag:20s} {t*1e6:8.1f} us  {N*K/t/1e12:5.2f} TB/s of weights  ({lib.mixdq_last_path().decode()})", flush=True)
    keep.clear()
lib.mixdq_debug_set_persist(1, 2); lib.mixdq_debug_set_persist_bn(0)
