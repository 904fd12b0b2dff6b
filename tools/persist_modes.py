"""What bounds the persistent tcgen05 kernel on tensor-bound GEMMs? Times a shape per tile width
and cluster size in three modes: 0 = normal, 1 = MMA issue skipped (pure TMA streaming through the
ring), 2 = TMA loads skipped (pure MMA issue + epilogue). Back-to-back launches in a CUDA graph,
weights rotated. Results of modes 1 / 2 are garbage by construction."""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()


def graph_time(fns, iters=5):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in fns:
            f()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for f in fns:
            f()
    g.replay(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters / len(fns) * 1e-3


# python tools/persist_modes.py [M,N,K ...]  — every (bn, cs) per shape; mode 1 (no MMA) beside mode 0
shapes = [(8192, 2560, 2560), (32768, 1280, 1280), (2048, 10240, 1280), (2048, 1280, 1280),
          (2048, 3840, 1280), (8192, 640, 640), (8192, 5120, 640), (2048, 1280, 5120)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for (M, N, K) in shapes:
    a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
    ws = [torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev) for _ in range(4)]
    z = torch.zeros(N, device=dev); o = torch.ones(N, device=dev); s1 = torch.tensor(1.0, device=dev)
    for bn in (128, 160, 256):
        if N % bn and bn == 160 and N % 32:
            continue
        for cs in (1, 2):
            line = f"M={M} N={N} K={K} bn={bn} cs={cs}:"
            for mode in (0, 1):
                lib.mixdq_debug_set_persist(2, cs); lib.mixdq_debug_set_persist_bn(bn)
                lib.mixdq_debug_set_mode(mode)
                outs = []
                t = graph_time([(lambda w=w: outs.append(ops.qlinear_w8_a8_ohalf(a, w, o, s1, s1, z, o, z, None)))
                                for w in ws] * 2)
                path = lib.mixdq_last_path().decode()
                line += f"  mode{mode} {t*1e6:7.1f} us {2.0*M*N*K/t/1e12:6.0f} TOP/s"
                del outs
            print(line, f"({path})", flush=True)
    # the one-tile-per-CTA kernel with its own heuristic, for reference
    lib.mixdq_debug_set_mode(0); lib.mixdq_debug_set_persist(0, 2); lib.mixdq_debug_set_persist_bn(0)
    outs = []
    t = graph_time([(lambda w=w: outs.append(ops.qlinear_w8_a8_ohalf(a, w, o, s1, s1, z, o, z, None)))
                    for w in ws] * 2)
    print(f"M={M} N={N} K={K} one-tile-per-CTA: {t*1e6:7.1f} us {2.0*M*N*K/t/1e12:6.0f} TOP/s "
          f"({lib.mixdq_last_path().decode()})", flush=True)
    del outs, ws, a
lib.mixdq_debug_set_mode(0); lib.mixdq_debug_set_persist(1, 2); lib.mixdq_debug_set_persist_bn(0)
