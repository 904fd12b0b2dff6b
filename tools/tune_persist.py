"""Per shape: one-tile-per-CTA kernel vs the persistent kernel at every tile width (CTA pairs),
same harness as the config-5 sweep (tools/layer_sweep.time_shape). Feeds the tile heuristic of
csrc/persist.cu::persist_pick_bn.   python tools/tune_persist.py [batches, default 1,8] [kind prefix]
(3x3 / stride 1 convolutions at width 160 run the HALO form unless MIXDQ_CONV_HALO=0)"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mixdq_b200 import _lib
from tools import layer_sweep as LS

dev = torch.device("cuda:0")
lib = _lib.load()
batches = [int(b) for b in (sys.argv[1] if len(sys.argv) > 1 else "1,8").split(",")]
only = sys.argv[2] if len(sys.argv) > 2 else ""        # kind prefix filter, e.g. conv3x3
for batch in batches:
    for kind, M, N, K, count in LS.APPENDIX_A:
        if M * batch < 256 or N < 64 or K % 16 or not kind.startswith(only):
            continue
        res = {}
        for tag, mode, bn in (("tile", 0, 0), ("p128", 2, 128), ("p160", 2, 160), ("p256", 2, 256)):
            lib.mixdq_debug_set_persist(mode, 2); lib.mixdq_debug_set_persist_bn(bn)
            try:
                t, nops, nbytes, path = LS.time_shape(kind, M, N, K, batch, dev)
            except RuntimeError as e:
                res[tag] = None
                continue
            res[tag] = (t * 1e6, path)
            torch.cuda.empty_cache()
        best = min((v[0], k) for k, v in res.items() if v)
        cells = "  ".join(f"{k} {v[0]:7.1f}{'*' if k == best[1] else ' '}" if v else f"{k}     n/a "
                          for k, v in res.items())
        print(f"b{batch} {kind:9s} M={M*batch:6d} N={N:5d} K={K:5d} x{count:3d} | {cells} | "
              f"{2.0*M*batch*N*K/best[0]/1e6:6.0f} TOP/s ({res['tile'][1] if res['tile'] else ''})", flush=True)
lib.mixdq_debug_set_persist(1, 2); lib.mixdq_debug_set_persist_bn(0)
