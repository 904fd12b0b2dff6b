"""Where do the dynamic quantiser kernels spend their time inside a dependent chain? A CUDA graph of
[GEMM(+residual) -> LayerNorm-quant -> GEMM -> plain quant] x L with per-CTA %globaltimer stamps in
both kernel families; every stamp is printed relative to the END of the preceding kernel
(mean/max over CTAs, ns)."""
import sys
import torch
sys.path.insert(0, ".")
from mixdq_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
L = 5
M, C = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (256, 1280)
MAXCTA = 1024

g = torch.Generator().manual_seed(0)
w = torch.randint(-127, 128, (C, C), dtype=torch.int8, generator=g).to(dev)
wsc = (0.001 + 0.01 * torch.rand(C, generator=g)).to(dev)
wsum = w.float().sum(1)
gamma = torch.ones(C, device=dev).half(); beta = torch.zeros(C, device=dev).half()
x0 = torch.randn(M, C, generator=g).half().to(dev)
tbuf = [torch.zeros(MAXCTA * 16, dtype=torch.int64, device=dev) for _ in range(2 * L)]
NQ = 4 * L + 2      # two kernels per quantisation (pass 1 + pass 2)
qbuf = torch.zeros(NQ * 1024 * 8, dtype=torch.int64, device=dev)
keep = []
ops.DYNAMIC_QUANT_CACHE = False


def body():
    x = x0
    q8, s, z = ops.layernorm_quantize_dynamic(x, gamma, beta, 1e-5)          # quant launch 0
    for i in range(L):
        lib.mixdq_debug_set_timing_buffer(tbuf[2 * i].data_ptr())
        x = ops.qlinear_dynamic_fused(q8, w, wsc, s, z, wsum, None, residual=x)
        q8, s, z = ops.layernorm_quantize_dynamic(x, gamma, beta, 1e-5)      # quant launch 1 + 2i
        lib.mixdq_debug_set_timing_buffer(tbuf[2 * i + 1].data_ptr())
        y = ops.qlinear_dynamic_fused(q8, w, wsc, s, z, wsum, None)
        q8, s, z = ops.quantize_per_tensor_dynamic(y)                        # quant launch 2 + 2i
        keep.extend([x, y, q8, s, z])
    lib.mixdq_debug_set_timing_buffer(None)


# ONE workspace, created eagerly: ops keys workspaces by stream, and one first created during
# capture would be zero-filled by a memset node at every replay (wiping the stamp pointer)
_ws = torch.zeros(lib.mixdq_quant_dynamic_ws_bytes(), dtype=torch.uint8, device=dev)
ops._dynamic_workspace = lambda device: _ws
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    body()
torch.cuda.current_stream().wait_stream(side)
graph = torch.cuda.CUDAGraph()
with torch.cuda.graph(graph):
    body()
# one workspace per (device, stream): the captured kernels use the capture stream's
wsps = [_ws.data_ptr()]
st = torch.cuda.current_stream().cuda_stream
for rep in range(4):
    for wsp in wsps:
        assert lib.mixdq_debug_set_quant_timing_buffer(wsp, qbuf.data_ptr(), st) == 0
    graph.replay()
torch.cuda.synchronize()
for wsp in wsps:
    lib.mixdq_debug_set_quant_timing_buffer(wsp, None, st)
torch.cuda.synchronize()

T = [b.cpu().view(-1, 16) for b in tbuf]
T = [t[t[:, 0] > 0] for t in T]
Qall = qbuf.cpu().view(NQ, 1024, 8)
Q = []
for k in range(NQ):
    t = Qall[k]
    Q.append(t[t[:, 0] > 0])
print("stamped CTAs per quantiser launch:", [q.shape[0] for q in Q])
qn = ["entry", "waited", "loaded", "published", "done"]
tn = ["entry", "setup", "tma0", "tmaN", "land0", "mmaN", "accrdy", "epi"]


def show(tag, stamps, ref, names):
    rel = (stamps - ref).float()
    print(f"{tag} ({stamps.shape[0]} CTAs): " +
          " ".join(f"{n}[{rel[:, j].mean():.0f}/{rel[:, j].max():.0f}]" for j, n in enumerate(names)))


# launch order: ln | per iteration: GEMM+res, ln, GEMM, quant — each quantisation = PER kernels
PER = int(sys.argv[3]) if len(sys.argv) > 3 else (2 if Q[1].shape[0] and Q[2 * L + 1].shape[0] else 1)
print("kernels per quantisation:", PER)
for i in range(1, L):
    g0_end = T[2 * i][:, 7].max()
    g1_end = T[2 * i + 1][:, 7].max()
    if PER == 2:
        base = 2 + 4 * i
        show(f"[{i}] ln pass1 after GEMM end", Q[base][:, :5], g0_end, qn)
        show(f"[{i}] ln pass2 after pass1 end", Q[base + 1][:, :5], Q[base][:, 4].max(), qn)
        show(f"[{i}] GEMM after ln pass2 end", T[2 * i + 1][:, :8], Q[base + 1][:, 4].max(), tn)
        show(f"[{i}] quant pass1 after GEMM end", Q[base + 2][:, :5], g1_end, qn)
        show(f"[{i}] quant pass2 after pass1 end", Q[base + 3][:, :5], Q[base + 2][:, 4].max(), qn)
        last = Q[base + 3]
    else:
        base = 1 + 2 * i
        show(f"[{i}] ln (one kernel) after GEMM end", Q[base][:, :5], g0_end, qn)
        show(f"[{i}] GEMM after ln end", T[2 * i + 1][:, :8], Q[base][:, 4].max(), tn)
        show(f"[{i}] quant (one kernel) after GEMM end", Q[base + 1][:, :5], g1_end, qn)
        last = Q[base + 1]
    if i + 1 < L:
        show(f"[{i}] next GEMM+res after quant end", T[2 * i + 2][:, :8], last[:, 4].max(), tn)
