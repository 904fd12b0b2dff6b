"""Which side bounds the tcgen05 mainloop at batch-1 shapes? Per-CTA %globaltimer stamps
(tools/phase_timing.py machinery) for a shape under three modes: 0 = normal, 1 = MMA issue skipped
(pure TMA streaming), 2 = TMA loads skipped (pure MMA issue). Prints ns relative to the end of the
previous launch, averaged over CTAs and launches; `main` = mmaN - land0."""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
names = ["entry", "setup", "tma0", "tmaN", "land0", "mmaN", "accrdy", "epi"]
MAXCTA = 4096
NL = 6


def run(M, N, K, bn, splits, mode):
    bufs = [torch.zeros(MAXCTA * 16, dtype=torch.int64, device=dev) for _ in range(NL)]
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(NL)]
    z = torch.zeros(N, device=dev); o = torch.ones(N, device=dev); s1 = torch.tensor(1.0, device=dev)
    lib.mixdq_debug_force_bn(bn); lib.mixdq_debug_force_splits(splits); lib.mixdq_debug_set_mode(mode)
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
    lib.mixdq_debug_force_bn(0); lib.mixdq_debug_force_splits(0); lib.mixdq_debug_set_mode(0)
    T = [b.cpu().view(-1, 16) for b in bufs]
    T = [t[t[:, 0] > 0] for t in T]
    ncta = T[0].shape[0]
    rel = []
    for i in range(1, NL):
        prev_end = T[i - 1][:, 7].max()
        rel.append((T[i] - prev_end).float())
    R = torch.stack(rel).mean(0)
    period = torch.stack([T[i][:, 7].max() - T[i - 1][:, 7].max() for i in range(1, NL)]).float().mean()
    line = f"M={M} N={N} K={K} BN={bn} S={splits} mode={mode} ctas={ncta} period={period:.0f} | "
    for j, n in enumerate(names):
        line += f"{n}[{R[:, j].mean():.0f}] "
    line += f"main[{(R[:, 5] - R[:, 4]).mean():.0f}] main_max[{(R[:, 5] - R[:, 4]).max():.0f}]"
    print(line, flush=True)


shapes = [(256, 1280, 1280), (256, 3840, 1280), (256, 10240, 1280), (1024, 640, 640)]
for (M, N, K) in shapes:
    for bn in (32, 64, 128, 256):
        if N == 10240 and bn < 128:
            continue
        for mode in (0, 1, 2):
            run(M, N, K, bn, 1, mode)
