"""In-kernel phase timestamps (globaltimer) of every CTA, inside a CUDA graph of back-to-back
launches (so launch gaps are the real ones)."""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
names = ["entry", "setup", "tma0", "tmaN", "land0", "mmaN", "accrdy", "epi", "phA", "clus", "sum", "staged"]
MAXCTA = 4096
NL = 6


def run(M, N, K, bn, splits):
    bufs = [torch.zeros(MAXCTA * 16, dtype=torch.int64, device=dev) for _ in range(NL)]
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(NL)]
    z = torch.zeros(N, device=dev); o = torch.ones(N, device=dev); s1 = torch.tensor(1.0, device=dev)
    lib.mixdq_debug_force_bn(bn); lib.mixdq_debug_force_splits(splits)
    outs = []

    def body():
        for i, w in enumerate(ws):
            lib.mixdq_debug_set_timing_buffer(bufs[i].data_ptr())
            outs.append(ops.qlinear_w8_a8_ohalf(a, w, o, s1, s1, z, o, z, None))
        lib.mixdq_debug_set_timing_buffer(None)
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        body()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    lib.mixdq_debug_force_bn(0); lib.mixdq_debug_force_splits(0)
    T = [b.cpu().view(-1, 16) for b in bufs]
    T = [t[t[:, 0] > 0] for t in T]
    ncta = T[0].shape[0]
    # launch i (i>=1): everything relative to the end of launch i-1 (max epi over CTAs)
    rel = []
    for i in range(1, NL):
        prev_end = T[i - 1][:, 7].max()
        t = (T[i] - prev_end).float()
        rel.append(t)
    R = torch.stack(rel).mean(0)          # [ncta, 8] averaged over launches
    per_launch = (torch.stack([T[i][:, 7].max() - T[i - 1][:, 7].max() for i in range(1, NL)]).float().mean())
    line = f"M={M} N={N} K={K} BN={bn} S={splits} ctas={ncta} period={per_launch:.0f}ns | "
    for j, n in enumerate(names):
        col = R[:, j]
        if j >= 8 and T[1][:, j].max() == 0:
            continue
        line += f"{n}[{col.mean():.0f}] "
    print(line, flush=True)


for cfg in [(256, 10240, 1280, 256, 1), (256, 1280, 1280, 64, 1), (256, 1280, 1280, 128, 4),
            (256, 1280, 1280, 128, 2), (256, 1280, 5120, 128, 8), (256, 1280, 5120, 256, 8)]:
    run(*cfg)
