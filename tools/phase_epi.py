"""Epilogue timeline of the batch-1 GEMMs: per-CTA %globaltimer stamps (thread 64 = first epilogue
warp) relative to the end of the previous launch in a graph of back-to-back launches.
  accrdy = accumulators complete, ld = first tcgen05.ld chunk in registers, staged = chunk
  dequantised + staged in smem, st = all stores of this warp issued, epi = CTA's epilogue done."""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
cols = {"entry": 0, "setup": 1, "tma0": 2, "tmaN": 3, "land0": 4, "mmaN": 5, "accrdy": 6,
        "ld": 12, "staged": 13, "st": 8, "epi": 7}
MAXCTA = 4096
NL = 6


def run(M, N, K, bn, dyn, res):
    bufs = [torch.zeros(MAXCTA * 16, dtype=torch.int64, device=dev) for _ in range(NL)]
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(NL)]
    z = torch.zeros(N, device=dev); o = torch.ones(N, device=dev); s1 = torch.tensor(1.0, device=dev)
    r = torch.randn(M, N, device=dev, dtype=torch.float16)
    lib.mixdq_debug_force_bn(bn)
    outs = []

    def body():
        for i, w in enumerate(ws):
            lib.mixdq_debug_set_timing_buffer(bufs[i].data_ptr())
            if dyn:
                outs.append(ops.qlinear_dynamic_fused(a, w, o, s1, s1, z, None, residual=r if res else None))
            else:
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
    lib.mixdq_debug_force_bn(0)
    T = [b.cpu().view(-1, 16) for b in bufs]
    T = [t[t[:, 0] > 0] for t in T]
    rel = []
    for i in range(1, NL):
        prev_end = T[i - 1][:, 7].max()
        rel.append((T[i] - prev_end).float())
    R = torch.stack(rel).mean(0)
    period = torch.stack([T[i][:, 7].max() - T[i - 1][:, 7].max() for i in range(1, NL)]).float().mean()
    line = f"M={M} N={N} K={K} BN={bn} dyn={dyn} res={res} ctas={T[0].shape[0]} period={period:.0f} | "
    for n, j in cols.items():
        line += f"{n}[{R[:, j].mean():.0f}] "
    print(line, flush=True)


for (M, N, K) in [(256, 1280, 1280), (256, 3840, 1280), (1024, 640, 640)]:
    for bn in (32, 64):
        run(M, N, K, bn, False, False)
        run(M, N, K, bn, True, True)
