#!/usr/bin/env python
"""bench.py — W8A8 SDXL-Turbo UNet step on B200 (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one UNet forward (1-step SDXL-Turbo sampling = one UNet evaluation) over one batch of
synthetic inputs: latents 64x64 (512x512 image), 77 text tokens, random-init weights of the named
architecture (no network for checkpoints). Workload at every N: BASELINE.json configs[1] per GPU
(SDXL-Turbo UNet, W8A8, batch 1) — weak scaling, batch-sharded data parallel, the final latents
all-gathered over NCCL inside the timed region.

Prints ONE JSON line (rank 0). `value` = images/s of the whole job with inputs resident in HBM and
the UNet replayed as one CUDA graph; `e2e` = the same through the public module call with pinned
HOST inputs, H2D copies and the D2H read of the latents inside the timed region; `roofline` = the
tcgen05 contraction kernel family (every GEMM / implicit-GEMM conv launch of one step re-issued
back to back as a graph, CUDA-event timed) as achieved algorithmic GB/s against the measured HBM
copy peak; `cpu_baseline` = the qdiff fake-quant oracle timed on the host cores on a bounded sample.

`--impl reference` times the reference's own CPU implementation of the path (qdiff fake-quant; the
reference is Python and /root/reference does not travel to the GPU box, so this is the oracle
port — the only other place oracle/ is executed) with all host threads.
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


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: qdiff fake-quant on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_fake_quant_sample(model_name: str, min_seconds: float, max_reps: int, threads: int,
                          min_reps: int = 2):
    """Time the fake-quant oracle on a bounded sample of the workload: the UNet's mid_block
    (2 resnets + its transformer blocks at the deepest resolution, batch 1). Returns
    (seconds per sample [list], MAC share of the sample, description)."""
    from mixdq_b200.unet import UNet2DConditionModel, MidBlock, sdxl_turbo_config, sd_turbo_config
    from oracle import unet_oracle as UO
    torch.set_num_threads(threads)
    cfg = {"sdxl-turbo": sdxl_turbo_config, "sd-turbo": sd_turbo_config}[model_name]()
    ch = cfg.block_out_channels[-1]
    with torch.device("meta"):
        mid = MidBlock(cfg, ch, cfg.block_out_channels[0] * 4, cfg.transformer_layers_per_block[-1])
    mid = mid.to_empty(device="cpu")
    fast_init_(mid, 0, "cpu")
    names = {n: 8 for n, m in mid.named_modules() if isinstance(m, (nn.Linear, nn.Conv2d))}
    UO.wrap_unet(mid, names, dict(names), {}, bos=False)
    res = cfg.sample_size // (2 ** (len(cfg.block_out_channels) - 1))
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, ch, res, res, generator=g)
    temb = torch.randn(1, cfg.block_out_channels[0] * 4, generator=g)
    ctx = torch.randn(1, 77, cfg.cross_attention_dim, generator=g)
    times = []
    with torch.no_grad():
        mid(x, temb, ctx)  # warm-up (allocator, thread pool)
        t_all = time.perf_counter()
        while len(times) < max_reps and (len(times) < min_reps
                                         or time.perf_counter() - t_all < min_seconds):
            t0 = time.perf_counter()
            mid(x, temb, ctx)
            times.append(time.perf_counter() - t0)
    macs = count_macs(model_name)
    share = sum(v for k, v in macs.items() if k.startswith("mid_block")) / sum(macs.values())
    desc = (f"qdiff fake-quant W8A8 of {model_name} mid_block (batch 1, {len(names)} of "
            f"{len(macs)} quantized layers, {share * 100:.2f}% of the step's MACs), fp32, "
            f"{threads} threads; img/s = MAC share / median seconds")
    return times, share, desc


def run_reference_arm(args, rank: int, world: int):
    """The reference's CPU implementation of the path, rank 0 only."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    reps = args.steps + args.warmup
    times, share, desc = cpu_fake_quant_sample(args.model, 0.0, reps, threads, min_reps=reps)
    timed = times[-args.steps:]
    sec = sum(timed) / len(timed)
    value = share / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / share * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.model} UNet W8A8 (qdiff fake-quant on CPU), 1 step, "
                               "512x512 (64x64 latent), batch 1", "sample": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def quantize_copy(unet_fp16, mode: str, w_config=None, a_config=None):
    from mixdq_b200 import mixdq
    q = copy.deepcopy(unet_fp16)
    names = [n for n, _ in q.quantizable_layers()]
    w_cfg = w_config or {n: 8 for n in names}
    a_cfg = a_config or {n: 8 for n in names}
    args = SimpleNamespace(w_config=w_cfg, a_config=a_cfg)
    ckpt = None
    if mode == "static":
        ckpt = synth_ckpt(q)
    mixdq.quantize_unet(q, args, ckpt=ckpt, bos=False, bos_dict=None)
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="sdxl-turbo", choices=["sdxl-turbo", "sd-turbo"])
    ap.add_argument("--batch", type=int, default=1, help="samples per GPU")
    ap.add_argument("--mode", default="dynamic", choices=["dynamic", "static"],
                    help="activation scales: dynamic per-tensor min-max (north star) or static ckpt")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp16", action="store_true")
    ap.add_argument("--profile-fp16", action="store_true",
                    help="with --profile-step: profile the FP16 baseline UNet instead")
    ap.add_argument("--profile-step", action="store_true",
                    help="run ONE eager quantized step between cudaProfilerStart/Stop and exit "
                         "(for `ncu --profile-from-start off`); prints no bench line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    from mixdq_b200 import dp
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        run_reference_arm(args, rank, int(os.environ.get("WORLD_SIZE", "1")))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback of the hot path)")
    rank, world, local = dp.init_distributed()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    from mixdq_b200 import _lib, ops
    _lib.load()

    B = args.batch
    unet16 = build_fp16_unet(args.model, device, seed=0)
    inputs = unet16.example_inputs(B, device, torch.float16, seed=1 + rank)
    hbm_peak, bf16_peak, peak_src = load_peaks()
    sampler = ClockSampler(local)

    if args.profile_step:
        qunet = unet16 if args.profile_fp16 else quantize_copy(unet16, args.mode)
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
    fp16_ms = None
    if not args.no_fp16:
        u16 = copy.deepcopy(unet16)
        g16, out16 = capture(u16, inputs)
        fp16_ms = time_region(g16.replay, args.steps, args.warmup, world, device)
        del u16, g16, out16
        torch.cuda.empty_cache()

    # ---- quantized UNet ----
    qunet = quantize_copy(unet16, args.mode)
    del unet16
    torch.cuda.empty_cache()
    n_layers = sum(1 for m in qunet.modules() if getattr(m, "valid_for_acceleration", False))

    # eager pass with the launch recorder on: launch inventory + kernel-family replay list
    with torch.no_grad():
        qunet(**inputs)
        c0 = ops.launch_count()
        rec = ops.start_recording()
        eager_out = qunet(**inputs)[0]
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
            dp.gather_latents(static_out[0], B * world, world)

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
            out = dp.gather_latents(out, B * world, world)[rank * B:(rank + 1) * B]
        host_out.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the caller reads the latents every step

    ms_e2e = time_region(step_e2e, args.steps, args.warmup, world, device)
    clocks = sampler.stop()

    # ---- roofline: the contraction kernel family of one step, re-issued back to back ----
    tc = [r for r in rec if r[0] in ("gemm", "gemm_geglu", "conv", "conv_split", "gemm_w4")]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for r in tc:
            r[3]()
    torch.cuda.current_stream().wait_stream(side)
    g_tc = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_tc):
        for r in tc:
            r[3]()
    tc_ms = time_region(g_tc.replay, max(args.steps, 10), 3, 1, device)
    tc_bytes = sum(r[1] for r in tc)
    tc_ops = sum(r[2] for r in tc)
    tc_launch = sum(r[5] for r in tc)
    achieved = tc_bytes / (tc_ms * 1e-3) / 1e9
    traffic = None
    prof = ROOT / "profiles" / "traffic.json"
    if prof.exists():
        try:
            traffic = json.loads(prof.read_text()).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
        "frac": achieved / hbm_peak, "traffic": traffic,
        "kernel": "tc_i8_kernel (tcgen05 int8 GEMM / implicit-GEMM conv family)",
        "launches_per_step": tc_launch, "avg_launch_us": tc_ms * 1e3 / max(tc_launch, 1),
        "algorithmic_bytes_per_launch": tc_bytes / max(tc_launch, 1),
        "family_ms_per_step": tc_ms, "share_of_step": tc_ms / ms,
        "tensor_tops": tc_ops / (tc_ms * 1e-3) / 1e12,
        "tensor_frac_of_2x_measured_bf16": tc_ops / (tc_ms * 1e-3) / 1e12 / (2 * bf16_peak),
        "peak_source": f"{peak_src} MEASURED_PEAKS.json hbm_gbs",
    }
    del g_tc

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        times, share, desc = cpu_fake_quant_sample(args.model, 12.0, 40, threads)
        med = sorted(times)[len(times) // 2]
        cpu_baseline = {"value": share / med, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": desc, "sample_seconds_median": med, "reps": len(times)}

    if rank == 0:
        total = B * world
        line = {
            "metric": METRIC, "value": total / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "s8",
            "data": "synthetic",
            "config": {
                "workload": f"{args.model} UNet W8A8 ({args.mode} per-tensor activation scales, "
                            f"all {n_layers} quantized layers), 1 step, 512x512 (64x64 latent), "
                            f"batch {B} per GPU, whole-UNet CUDA graph",
                "global_batch": total, "parallelism": f"dp{world}",
                "l2": "inputs larger than L2 (2.57 GB of int8 weights streamed per step vs 126 MB L2)",
            },
            "e2e": {"value": total / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches_per_step * args.steps,
            "gpu_launches_per_step": launches_per_step,
            "launch_families": {k: {"calls": v[0], "kernels": v[3]} for k, v in fam.items()},
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "clocks": clocks,
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
