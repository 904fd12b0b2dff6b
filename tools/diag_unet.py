"""Diagnose a whole-UNet mismatch between the GPU W8A8 path and the CPU fake-quant oracle.

For every quantized leaf, in execution order:
  free   : relative error of the GPU layer OUTPUT inside the free-running GPU UNet vs the oracle's
           output inside the free-running oracle UNet (shows where the trajectories diverge);
  forced : the GPU layer fed the ORACLE's input (rounded to fp16) vs the oracle layer's output
           (teacher-forced: isolates each layer; a bug shows up as one bad layer).
Runs on the GPU box (python tools/diag_unet.py [--bos]); not part of the product path.
"""
import argparse
import copy
import sys
from pathlib import Path
from types import SimpleNamespace

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item(), \
        torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bos", action="store_true")
    ap.add_argument("--simt", action="store_true")
    args = ap.parse_args()
    from mixdq_b200 import mixdq, _lib
    from mixdq_b200.quantize import derive_up_block_splits
    from mixdq_b200.unet import build_unet
    from oracle import unet_oracle as UO
    dev = torch.device("cuda:0")
    if args.simt:
        _lib.load().mixdq_force_simt(1)
    unet = build_unet("tiny", seed=3).half()
    names = [n for n, _ in unet.quantizable_layers()]
    protect = ("conv_in", "conv_out")
    w_bits = {n: 8 for n in names}
    a_bits = {n: 8 for n in names if n not in protect}
    ref_unet = UO.wrap_unet(copy.deepcopy(unet).float(), w_bits, a_bits,
                            derive_up_block_splits(unet), bos=args.bos)
    fp_unet = copy.deepcopy(unet).to(dev)
    inputs = unet.example_inputs(2, "cpu", torch.float16, seed=1)
    bos_dict = mixdq.compute_bos_dict(unet, inputs["encoder_hidden_states"]) if args.bos else None
    mixdq.quantize_unet(unet, SimpleNamespace(w_config=w_bits, a_config=a_bits), ckpt=None,
                        bos=args.bos, bos_dict=bos_dict)
    unet = unet.to(dev).to(memory_format=torch.channels_last)

    rec_ref, rec_gpu, order = {}, {}, []

    def hook(store, name, keep_order=False):
        def f(m, inp, out):
            store[name] = (inp[0].detach().cpu(), out.detach().cpu())
            if keep_order:
                order.append(name)
        return f
    for n in names:
        ref_unet.get_submodule(n).register_forward_hook(hook(rec_ref, n, True))
        unet.get_submodule(n).register_forward_hook(hook(rec_gpu, n))
    with torch.no_grad():
        got = unet(**{k: v.to(dev) for k, v in inputs.items()})[0]
        ref = ref_unet(**{k: (v.float() if v.is_floating_point() else v)
                          for k, v in inputs.items()})[0]
        fp = fp_unet(**{k: v.to(dev) for k, v in inputs.items()})[0]
    print("final  gpu-w8a8 vs oracle-fakequant:", rel(got, ref))
    print("final  gpu-fp16 (no quant) vs oracle-fakequant:", rel(fp, ref))
    print(f"{'layer':70s} {'free_in':>9s} {'free_out':>9s} {'forced':>9s} cos_forced")
    with torch.no_grad():
        for n in order:
            xi, yo = rec_ref[n]
            gi, go = rec_gpu[n]
            m = unet.get_submodule(n)
            xin = xi.half().to(dev)
            if xin.dim() == 4:
                xin = xin.contiguous(memory_format=torch.channels_last)
            yf = m(xin)
            e_in, _ = rel(gi, xi)
            e_out, _ = rel(go, yo)
            e_f, c_f = rel(yf, yo)
            flag = "  <<<" if e_f > 1e-2 else ""
            print(f"{n:70s} {e_in:9.2e} {e_out:9.2e} {e_f:9.2e} {c_f:.6f}{flag}")


if __name__ == "__main__":
    main()
