#!/usr/bin/env python
"""bench.py — the quantized-UNet hot path on B200 (BASELINE.json metric), one process per GPU.

  python bench.py [--config 2] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one UNet forward (1-step Turbo sampling = one UNet evaluation) over one batch of
synthetic inputs: latents 64x64 (512x512 image), 77 text tokens, random-init weights of the named
architecture (no network for checkpoints). `--config` selects one of BASELINE.json's configurations:

  1  qdiff fake-quant W8A8 SDXL-Turbo UNet, batch 1, on the CPU (= `--impl reference`)
  2  SDXL-Turbo UNet W8A8, batch 1 per GPU, vs the FP16 UNet            [default; weak scaling]
  3  SDXL-Turbo UNet mixed precision (weight_5.02.yaml: W4/W8; act_8.00.yaml), batch 8 per GPU
  4  SD-Turbo (SD2.1 UNet) W8A8, GLOBAL batch 64 sharded over the GPUs  [strong scaling]
  5  per-layer sweep of the SDXL QuantLinear / QuantConv2d shapes: INT8 TOP/s vs the roofline

Prints ONE JSON line (rank 0). `value` = images/s of the whole job with inputs resident in HBM and
the UNet replayed as one CUDA graph; `e2e` = the same through the public module call with pinned
HOST inputs, H2D copies and the D2H read of the latents inside the timed region; `roofline` = the
tcgen05 contraction kernel family (every GEMM / implicit-GEMM conv launch of one step re-issued
back to back as a graph, CUDA-event timed) as achieved algorithmic GB/s against the measured HBM
copy peak; `cpu_baseline` = the qdiff fake-quant oracle (whole UNet, batch 1) timed on the host
cores; `memory` = static / dynamic / peak device memory beside the reference's published numbers.

`--impl reference` times the reference's own CPU implementation of the path (qdiff fake-quant of
the WHOLE UNet, batch 1; the reference is Python and /root/reference does not travel to the GPU
box, so this is the oracle port — the only other place oracle/ is executed) with all host threads.
"""
from __future__ import annotations

import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path
from types import SimpleNamespace

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

METRIC = "W8A8 SDXL-Turbo UNet images/s (1-step 512x512; ms_per_step = UNet ms/step)"
UNIT = "img/s"

# BASELINE.json configs 2-4 (1 = the CPU arm, 5 = the per-layer sweep)
CONFIGS = {
    2: dict(model="sdxl-turbo", batch=1, global_batch=None, w_config=None, a_config=None,
            mode="dynamic", scaling="weak",
            what="SDXL-Turbo UNet W8A8 (all 794 layers), batch 1 per GPU"),
    3: dict(model="sdxl-turbo", batch=8, global_batch=None, w_config="weight/weight_5.02.yaml",
            a_config="act/act_8.00.yaml", mode="dynamic", scaling="weak",
            what="SDXL-Turbo UNet mixed precision weight_5.02.yaml (401 W8 / 246 W4 / 147 W2->W4) "
                 "+ act_8.00.yaml (9 layers fp16), batch 8 per GPU"),
    4: dict(model="sd-turbo", batch=None, global_batch=64, w_config=None, a_config=None,
            mode="static", scaling="strong",
            what="SD-Turbo (SD2.1) UNet W8A8, static scales, global batch 64 sharded over the GPUs"),
}


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
def fast_init_(model: nn.Module, seed: int, device) -> None:
    """PyTorch-default-like init (uniform +-1/sqrt(fan_in)) done in place on `device`; the default
    constructors would spend a minute initialising 2.6 G parameters on the CPU."""
    g = torch.Generator(device=device).manual_seed(seed)
    for m in model.modules():
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            fan_in = m.weight[0].numel()
            bound = 1.0 / fan_in ** 0.5
            m.weight.data.uniform_(-bound, bound, generator=g)
            if m.bias is not None:
                m.bias.data.uniform_(-bound, bound, generator=g)
        elif isinstance(m, (nn.GroupNorm, nn.LayerNorm)):
            m.weight.data.fill_(1.0)
            m.bias.data.zero_()


def tiled_init_cpu_(model: nn.Module, seed: int) -> None:
    """CPU variant for the timing-only fake-quant arm: one 16 M-element uniform block tiled over
    every weight (memcpy speed; `uniform_` over 2.6 G fp32 elements takes ~35 s single-threaded).
    Same value distribution per layer as fast_init_, which is all the timing depends on."""
    g = torch.Generator().manual_seed(seed)
    block = torch.empty(1 << 24).uniform_(-1.0, 1.0, generator=g)
    for m in model.modules():
        if isinstance(m, (nn.Linear, nn.Conv2d)):
            bound = 1.0 / m.weight[0].numel() ** 0.5
            for t in (m.weight, m.bias):
                if t is None:
                    continue
                flat = t.data.view(-1)
                for o in range(0, flat.numel(), block.numel()):
                    n = min(block.numel(), flat.numel() - o)
                    torch.mul(block[:n], bound, out=flat[o:o + n])
        elif isinstance(m, (nn.GroupNorm, nn.LayerNorm)):
            m.weight.data.fill_(1.0)
            m.bias.data.zero_()


def build_fp16_unet(name: str, device, seed: int = 0):
    from mixdq_b200.unet import build_unet
    with torch.device("meta"):
        unet = build_unet(name)
    unet = unet.to_empty(device=device).half()
    fast_init_(unet, seed, device)
    return unet.to(memory_format=torch.channels_last).eval()


def count_macs(name: str, batch: int = 1):
    from mixdq_b200.unet import build_unet
    with torch.device("meta"):
        u = build_unet(name)
    macs = {}

    def hook(n):
        def f(m, inp, out):
            k = m.in_features if isinstance(m, nn.Linear) else \
                m.in_channels * m.kernel_size[0] * m.kernel_size[1]
            macs[n] = macs.get(n, 0) + out.numel() * k
        return f
    for n, m in u.quantizable_layers():
        m.register_forward_hook(hook(n))
    inp = {k: v.to("meta") for k, v in u.example_inputs(batch, "cpu", torch.float32).items()}
    u(**inp)
    return macs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def time_region(fn, steps: int, warmup: int, world: int, device):
    """W untimed + exactly K timed calls of fn, barrier + synchronize on both sides, CUDA events on
    the launching stream, MAX over ranks. Returns ms per step."""
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return ms / steps


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d["hbm_gbs"], d["bf16_tflops"], "measured"
    return 6650.0, 1590.0, "fallback"


def bit_configs(unet, w_config, a_config):
    """(w_bits, a_bits) dicts of a configuration; None = every quantizable layer at 8 bit."""
    from mixdq_b200 import mixdq
    names = [n for n, _ in unet.quantizable_layers()]
    w = mixdq.load_bit_config(w_config) if w_config else {n: 8 for n in names}
    a = mixdq.load_bit_config(a_config) if a_config else {n: 8 for n in names}
    return w, a


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: qdiff fake-quant of the WHOLE UNet on the host cores
# ------------------------------------------------------------------------------------------------
class CpuFakeQuantUNet:
    """the reference's fake-quant path (quantizers initialised from the input of
    the same forward = dynamic min-max quantisation; weights re-quantised every forward, as
    QuantLayer.forward does), fp32, batch 1, whole UNet, all host threads."""

    def __init__(self, model_name: str, w_config=None, a_config=None, threads=None):
        from mixdq_b200.quantize import derive_up_block_splits
        from mixdq_b200.unet import build_unet
        from oracle import unet_oracle as UO
        self.threads = threads or os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        with torch.device("meta"):
            unet = build_unet(model_name)
        unet = unet.to_empty(device="cpu")
        tiled_init_cpu_(unet, 0)
        w_bits, a_bits = bit_configs(unet, w_config, a_config)
        self.n_layers = len(w_bits)
        UO.wrap_unet(unet, w_bits, a_bits, derive_up_block_splits(unet), bos=False)
        self.unet = unet.eval()
        self.inputs = unet.example_inputs(1, "cpu", torch.float32, seed=1)
        self.desc = (f"qdiff fake-quant of the WHOLE {model_name} UNet ({self.n_layers} quantized "
                     f"layers, batch 1, 1 step, 64x64 latent), fp32, {self.threads} threads")

    def step(self) -> float:
        t0 = time.perf_counter()
        with torch.no_grad():
            self.unet(**self.inputs)
        return time.perf_counter() - t0


def run_reference_arm(args, cfg, rank: int):
    """The reference's CPU implementation of the path, rank 0 only: W warm-up + exactly K timed
    whole-UNet forwards at batch 1 (for configs whose per-GPU batch is larger, one step of this
    arm is the bounded sample 'one image of the batch')."""
    if rank != 0:
        return
    cpu = CpuFakeQuantUNet(cfg["model"], cfg["w_config"], cfg["a_config"])
    for _ in range(args.warmup):
        cpu.step()
    times = [cpu.step() for _ in range(args.steps)]
    sec = sum(times) / len(times)
    value = 1.0 / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{cfg['what']} — reference arm: {cpu.desc}",
                   "config_id": args.config, "sample": cpu.desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.threads, "kind": "port",
                         "sample": cpu.desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def quantize_copy(unet_fp16, mode: str, w_config=None, a_config=None, fuse=None):
    from mixdq_b200 import mixdq
    q = copy.deepcopy(unet_fp16)
    w_cfg, a_cfg = bit_configs(q, w_config, a_config)
    args = SimpleNamespace(w_config=w_cfg, a_config=a_cfg)
    ckpt = synth_ckpt(q) if mode == "static" else None
    mixdq.quantize_unet(q, args, ckpt=ckpt, bos=False, bos_dict=None, fuse=fuse)
    return q.to(memory_format=torch.channels_last).eval()


def synth_ckpt(unet):
    """A PTQ checkpoint in the reference's kernel format with min-max weight scales of the
    random-init weights and plausible activation parameters (sigma~1 activations)."""
    from mixdq_b200.nn.utils import minmax_weight_scales
    from mixdq_b200.quantize import derive_up_block_splits
    splits = derive_up_block_splits(unet)
    ck = {}
    for n, m in unet.quantizable_layers():
        def entry(w):
            d = torch.stack([minmax_weight_scales(w, b) for b in (2, 4, 8)]).half().cpu()
            return {"delta_list": d, "zero_point_list": torch.zeros_like(d)}
        act = {"delta_list": torch.tensor([2.7, 0.55, 0.0323]).half(),
               "zero_point_list": torch.tensor([2.0, 8.0, 128.0]).half()}
        s = splits.get(n, 0)
        if s:
            ck[n + ".weight_quantizer"] = entry(m.weight[:, :s])
            ck[n + ".weight_quantizer_0"] = entry(m.weight[:, s:])
            ck[n + ".act_quantizer_0"] = act
        else:
            ck[n + ".weight_quantizer"] = entry(m.weight)
        ck[n + ".act_quantizer"] = act
    return ck


def capture(unet, inputs):
    """whole-UNet CUDA graph through the public API (mixdq.cuda_graph_opt)."""
    from mixdq_b200 import mixdq
    mixdq.cuda_graph_opt(unet)
    with torch.no_grad():
        unet(**inputs)
    (static_in, graph, static_out, _stream) = next(iter(unet.forward._cached.values()))
    return graph, static_out


def module_bytes(m: nn.Module) -> int:
    seen, total = set(), 0
    for t in list(m.parameters()) + list(m.buffers()):
        st = t.untyped_storage()
        if st.data_ptr() in seen:
            continue
        seen.add(st.data_ptr())
        total += st.nbytes()
    return total


def measure_memory(unet, inputs, device):
    """static = parameter/buffer storage of the UNet; dynamic = peak extra device memory of one
    eager forward (what the reference's kernels/README.md:76-91 printout calls dynamic)."""
    torch.cuda.synchronize(device)
    torch.cuda.empty_cache()
    static = module_bytes(unet)
    base = torch.cuda.memory_allocated(device)
    torch.cuda.reset_peak_memory_stats(device)
    with torch.no_grad():
        unet(**inputs)
    torch.cuda.synchronize(device)
    dyn = torch.cuda.max_memory_allocated(device) - base
    mb = 1.0 / 2 ** 20
    return {"static_mb": static * mb, "dynamic_mb": dyn * mb, "peak_mb": (static + dyn) * mb}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5],
                    help="BASELINE.json configuration (see the module docstring)")
    ap.add_argument("--model", default=None, choices=["sdxl-turbo", "sd-turbo"])
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU")
    ap.add_argument("--global-batch", type=int, default=None,
                    help="strong scaling: total samples, sharded over the GPUs")
    ap.add_argument("--mode", default=None, choices=["dynamic", "static"],
                    help="activation scales: dynamic per-tensor min-max (north star) or static ckpt")
    ap.add_argument("--w-config", default=None, help="weight bit config (packaged id or YAML path)")
    ap.add_argument("--a-config", default=None, help="activation bit config")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp16", action="store_true")
    ap.add_argument("--no-fuse", action="store_true", help="leaf-by-leaf module swap only")
    ap.add_argument("--no-static", action="store_true",
                    help="config 2: skip the extra static-scale (PTQ checkpoint) timing")
    ap.add_argument("--sweep-out", default=None, help="config 5: also write the table to this file")
    ap.add_argument("--profile-fp16", action="store_true",
                    help="with --profile-step: profile the FP16 baseline UNet instead")
    ap.add_argument("--profile-step", action="store_true",
                    help="run ONE eager quantized step between cudaProfilerStart/Stop and exit "
                         "(for `ncu --profile-from-start off`); prints no bench line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    cfg = dict(CONFIGS[args.config if args.config in CONFIGS else 2])
    for k in ("model", "batch", "mode", "w_config", "a_config"):
        if getattr(args, k) is not None:
            cfg[k] = getattr(args, k)
            cfg["what"] += f" [{k}={getattr(args, k)}]"
    if args.global_batch is not None:
        cfg["global_batch"], cfg["scaling"] = args.global_batch, "strong"

    if args.impl == "reference" or args.config == 1:
        run_reference_arm(args, cfg, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback of the hot path)")
    from mixdq_b200 import dp
    rank, world, local = dp.init_distributed()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    from mixdq_b200 import _lib, ops
    _lib.load()

    if args.config == 5:
        from tools import layer_sweep
        layer_sweep.run(args, rank, world, device)
        return

    if cfg["global_batch"] is not None:          # strong scaling: shard the global batch
        lo, hi = dp.shard_bounds(cfg["global_batch"], world, rank)
        B, total = hi - lo, cfg["global_batch"]
    else:
        B, total = cfg["batch"], cfg["batch"] * world
    model = cfg["model"]
    unet16 = build_fp16_unet(model, device, seed=0)
    inputs = unet16.example_inputs(B, device, torch.float16, seed=1 + rank)
    hbm_peak, bf16_peak, peak_src = load_peaks()
    sampler = ClockSampler(local)
    fuse = False if args.no_fuse else None

    if args.profile_step:
        qunet = unet16 if args.profile_fp16 else quantize_copy(unet16, cfg["mode"], cfg["w_config"],
                                                               cfg["a_config"], fuse)
        with torch.no_grad():
            for _ in range(2):
                qunet(**inputs)
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            qunet(**inputs)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        return

    # ---- FP16 baseline: same skeleton, cuBLAS/cuDNN, whole-UNet CUDA graph ----
    fp16_ms, mem16 = None, None
    if not args.no_fp16:
        u16 = copy.deepcopy(unet16)
        mem16 = measure_memory(u16, inputs, device)
        g16, out16 = capture(u16, inputs)
        fp16_ms = time_region(g16.replay, args.steps, args.warmup, world, device)
        del u16, g16, out16
        torch.cuda.empty_cache()

    # ---- quantized UNet ----
    qunet = quantize_copy(unet16, cfg["mode"], cfg["w_config"], cfg["a_config"], fuse)
    want_static = (args.config == 2 and cfg["mode"] == "dynamic" and not args.no_static)
    if not want_static:
        del unet16
    torch.cuda.empty_cache()
    kinds = {}
    for m in qunet.modules():
        if hasattr(m, "valid_for_acceleration"):
            k = m._get_name()
            kinds[k] = kinds.get(k, 0) + 1
    n_layers = sum(v for k, v in kinds.items() if "Fallback" not in k)
    mem8 = measure_memory(qunet, inputs, device)

    # eager pass with the launch recorder on: launch inventory + kernel-family replay list
    with torch.no_grad():
        qunet(**inputs)
        c0 = ops.launch_count()
        rec = ops.start_recording()
        qunet(**inputs)
        ops.stop_recording()
        launches_per_step = ops.launch_count() - c0
    fam = {}
    for family, nbytes, nops, _, _, k in rec:
        f = fam.setdefault(family, [0, 0, 0, 0])
        f[0] += 1; f[1] += nbytes; f[2] += nops; f[3] += k

    # whole-UNet graph through the public API
    graph, static_out = capture(qunet, inputs)

    def step_device():
        graph.replay()
        if world > 1:
            dp.gather_latents(static_out[0], total, world)

    sampler.start()
    torch.cuda.profiler.start()      # `ncu --profile-from-start off` captures exactly this region
    ms = time_region(step_device, args.steps, args.warmup, world, device)
    torch.cuda.profiler.stop()

    # ---- e2e: pinned host inputs -> H2D -> public forward (graph) -> D2H latents ----
    host_in = {k: (v.cpu().pin_memory() if torch.is_tensor(v) else v) for k, v in inputs.items()}
    host_out = torch.empty(static_out[0].shape, dtype=torch.float16).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host_in.values() if torch.is_tensor(v))
    d2h = host_out.numel() * host_out.element_size()

    def step_e2e():
        dev_in = {k: v.to(device, non_blocking=True) for k, v in host_in.items()}
        dev_in["sample"] = dev_in["sample"].contiguous(memory_format=torch.channels_last)
        with torch.no_grad():
            out = qunet(**dev_in)[0]
        if world > 1:
            lo = dp.shard_bounds(total, world, rank)[0]
            out = dp.gather_latents(out, total, world)[lo:lo + B]
        host_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the caller reads the latents every step

    ms_e2e = time_region(step_e2e, args.steps, args.warmup, world, device)
    clocks = sampler.stop()

    # ---- roofline: the contraction kernel family of one step, re-issued back to back ----
    tc_fams = ("gemm", "gemm_geglu", "conv", "conv_split", "gemm_w4", "gemm_geglu_w4", "conv_w4")
    tc = [r for r in rec if r[0] in tc_fams]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.prepare_stream(device)           # this stream's split-K workspace, outside the capture
        for r in tc:
            r[3]()
    torch.cuda.current_stream().wait_stream(side)
    g_tc = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_tc, stream=side):
        for r in tc:
            r[3]()
    tc_ms = time_region(g_tc.replay, max(args.steps, 10), 3, 1, device)
    tc_bytes = sum(r[1] for r in tc)
    tc_ops = sum(r[2] for r in tc)
    tc_launch = sum(r[5] for r in tc)
    achieved = tc_bytes / (tc_ms * 1e-3) / 1e9
    traffic = None
    prof = ROOT / "profiles" / "traffic.json"
    if prof.exists() and args.config == 2:
        try:
            traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    tops = tc_ops / (tc_ms * 1e-3) / 1e12
    # which side of the roofline binds the family at this batch (ridge ~ 500 op/B)
    ai = tc_ops / max(tc_bytes, 1)
    tensor_bound = ai > (2 * bf16_peak * 1e12) / (hbm_peak * 1e9)
    roofline = {
        "bound": "tensor" if tensor_bound else "hbm",
        "achieved": tops if tensor_bound else achieved,
        "peak": 2 * bf16_peak if tensor_bound else hbm_peak,
        "unit": "TFLOP/s" if tensor_bound else "GB/s",
        "frac": tops / (2 * bf16_peak) if tensor_bound else achieved / hbm_peak,
        "traffic": traffic,
        "kernel": "tc_i8_kernel / tc_i8_persist_kernel (tcgen05 int8 GEMM / implicit-GEMM conv family)",
        "launches_per_step": tc_launch, "avg_launch_us": tc_ms * 1e3 / max(tc_launch, 1),
        "algorithmic_bytes_per_launch": tc_bytes / max(tc_launch, 1),
        "algorithmic_ops_per_byte": ai,
        "family_ms_per_step": tc_ms, "share_of_step": tc_ms / ms,
        "hbm_gbs": achieved, "hbm_frac": achieved / hbm_peak,
        "tensor_tops": tops, "tensor_frac_of_2x_measured_bf16": tops / (2 * bf16_peak),
        "peak_source": f"{peak_src} MEASURED_PEAKS.json "
                       + ("2 x bf16_tflops (no INT8 peak is measured)" if tensor_bound else "hbm_gbs"),
    }
    del g_tc

    # ---- the same UNet with STATIC (PTQ-checkpoint) activation scales: the mode the reference's
    #      own kernels ship with (nn/Linear.py:154-176); reported beside the dynamic headline ----
    static_scales = None
    if want_static:
        sunet = quantize_copy(unet16, "static", cfg["w_config"], cfg["a_config"], fuse)
        del unet16
        with torch.no_grad():
            sunet(**inputs)
            c0 = ops.launch_count()
            sunet(**inputs)
            s_launches = ops.launch_count() - c0
        g_s, out_s = capture(sunet, inputs)

        def step_static():
            g_s.replay()
            if world > 1:
                dp.gather_latents(out_s[0], total, world)
        ms_s = time_region(step_static, args.steps, args.warmup, world, device)
        static_scales = {
            "ms_per_step": ms_s, "img_per_s": total / (ms_s * 1e-3),
            "gpu_launches_per_step": s_launches,
            "speedup_over_fp16": None if fp16_ms is None else fp16_ms / ms_s,
            "what": "same UNet, same kernels, static per-tensor activation scales from a synthetic "
                    "PTQ checkpoint in the reference's kernel format (the reference extension's own "
                    "mode): LayerNorm / GroupNorm / GEGLU producers emit int8 directly, no min/max "
                    "and no quantise pass behind them"}
        del g_s, out_s, sunet
        torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N = 1 only): whole UNet, batch 1, 1 warm-up + 3 timed ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = CpuFakeQuantUNet(model, cfg["w_config"], cfg["a_config"])
        cpu.step()
        times = sorted(cpu.step() for _ in range(3))
        med = times[1]
        cpu_baseline = {"value": 1.0 / med, "unit": UNIT, "cores": cpu.threads, "kind": "port",
                        "sample": cpu.desc + "; 1 warm-up + 3 timed forwards, median"
                        + ("" if B == 1 else f" (one image of the batch of {B})"),
                        "sample_seconds_median": med, "reps": 3}
        del cpu

    if rank == 0:
        mb = 1.0 / 2 ** 20
        line = {
            "metric": METRIC, "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
            "dtype": "s8", "data": "synthetic",
            "config": {
                "workload": f"{cfg['what']}; {cfg['mode']} per-tensor activation scales, "
                            f"{n_layers} layers on the int8 kernels, 1 step, 512x512 (64x64 latent), "
                            f"batch {B} on this GPU, whole-UNet CUDA graph",
                "config_id": args.config, "model": model, "global_batch": total,
                "parallelism": f"dp{world}", "layer_kinds": kinds,
                "l2": f"inputs larger than L2 ({mem8['static_mb']:.0f} MB of quantized weights "
                      "streamed per step vs 126 MB L2)",
            },
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "launch_families": {k: {"calls": v[0], "kernels": v[3]} for k, v in fam.items()},
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
            "memory": {
                "quantized": mem8, "fp16": mem16,
                "static_ratio_fp16_over_quantized":
                    None if mem16 is None else mem16["static_mb"] / mem8["static_mb"],
                "reference_published_mb": {
                    "what": "SDXL-Turbo UNet, batch 1 (reference README.md:41-45, "
                            "kernels/README.md:81-91; GPU unspecified)",
                    "fp16": {"static": 4998.0, "dynamic": 240.88, "peak": 5239.0},
                    "w8a8": {"static": 2575.32, "dynamic": 55.77, "peak": 2631.10}},
            },
            "static_scales": static_scales,
            "fp16_baseline": None if fp16_ms is None else {
                "ms_per_step": fp16_ms, "img_per_s": total / (fp16_ms * 1e-3),
                "speedup_w8a8_over_fp16": fp16_ms / ms,
                "what": "same UNet skeleton in fp16 (cuBLAS/cuDNN/SDPA via PyTorch), CUDA graph"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
