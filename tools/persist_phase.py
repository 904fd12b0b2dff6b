"""Steady-state timeline of the persistent tcgen05 kernel (tile 2 / 3 of every CTA), from the
%globaltimer stamps of tc_persist.cuh (TP_DBG): python tools/persist_phase.py M,N,K [bn] [cs] [mode] [tile]"""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
M, N, K = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "8192,5120,640").split(","))
bn = int(sys.argv[2]) if len(sys.argv) > 2 else 256
cs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
tile = int(sys.argv[5]) if len(sys.argv) > 5 else 2      # which tile of every CTA is stamped
lib.mixdq_debug_set_persist(2, cs); lib.mixdq_debug_set_persist_bn(bn); lib.mixdq_debug_set_mode(mode | (tile << 4))
a = torch.randint(-128, 128, (M, K), dtype=torch.int8, device=dev)
w = torch.randint(-128, 128, (N, K), dtype=torch.int8, device=dev)
z = torch.zeros(N, device=dev); o = torch.ones(N, device=dev); s1 = torch.tensor(1.0, device=dev)
buf = torch.zeros(4096 * 16, dtype=torch.int64, device=dev)
for _ in range(3):
    ops.qlinear_w8_a8_ohalf(a, w, o, s1, s1, z, o, z, None)
lib.mixdq_debug_set_timing_buffer(buf.data_ptr())
ops.qlinear_w8_a8_ohalf(a, w, o, s1, s1, z, o, z, None)
lib.mixdq_debug_set_timing_buffer(None)
torch.cuda.synchronize()
T = buf.cpu().view(-1, 16)
T = T[T[:, 5] > 0].double()
names = {11: "kernel entry", 12: "setup done", 13: "dep wait passed", 14: "roles done",
         0: "epi wait", 1: "acc ready", 2: "chunk ld", 3: "chunk staged", 4: "chunk stored",
         5: "tile2 drained", 6: "tile3 drained", 7: "mma tile2 wait", 10: "mma tile2 landed",
         8: "mma tile2 issued", 9: "mma tile3 wait"}
base = T[:, 1]
print(f"M={M} N={N} K={K} bn={bn} cs={cs} mode={mode}: {T.shape[0]} CTAs stamped; ns relative to 'acc ready' of tile {tile}")
for k, n in names.items():
    col = T[:, k]
    ok = col > 0
    if ok.any():
        d = (col[ok] - base[ok])
        print(f"  {n:18s} mean {d.mean():9.0f}  min {d.min():9.0f}  max {d.max():9.0f}")
ok = T[:, 6] > 0
if ok.any():
    print(f"  epilogue period (next tile drained - this tile drained): {(T[ok, 6] - T[ok, 5]).mean():.0f} ns")
