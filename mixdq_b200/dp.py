"""Batch-sharded data parallelism for the quantized UNet (new: the reference is single-GPU).

One process per GPU (torchrun), every rank holds a replica of the int8 weights (2.57 GB SDXL /
0.87 GB SD2.1 — replication is free on 180 GB parts), the global batch of prompts/latents is split
contiguously across ranks, and there is NO per-layer collective: with per-tensor activation scales
each sample's UNet step is independent. The only exchange is one all-gather of the final latents
(32 KiB per sample) over NCCL/NVLink per step batch; gloo is used for the CPU tests of this logic.

Note on dynamic activation scales: min/max is taken over the LOCAL shard, so codes can differ from
a single-GPU run of the whole batch; static (checkpoint) scales are bit-identical under any split.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment. Returns (rank, world, local)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_bounds(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split; the first `global_batch % world` ranks take one extra sample."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_inputs(inputs: Dict[str, torch.Tensor], global_batch: int, world: int, rank: int
                 ) -> Dict[str, torch.Tensor]:
    """Slice every per-sample tensor (leading dim == global_batch) to this rank's shard."""
    lo, hi = shard_bounds(global_batch, world, rank)
    out = {}
    for k, v in inputs.items():
        if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == global_batch:
            out[k] = v[lo:hi]
        else:
            out[k] = v
    return out


def gather_latents(local: torch.Tensor, global_batch: int, world: int) -> torch.Tensor:
    """All-gather the per-rank latents into [global_batch, ...] on every rank (ragged shards are
    padded to the largest shard for the collective and trimmed afterwards)."""
    if world == 1:
        return local
    sizes = [shard_bounds(global_batch, world, r) for r in range(world)]
    max_n = max(hi - lo for lo, hi in sizes)
    local = local.contiguous()
    if local.shape[0] < max_n:
        pad = torch.zeros((max_n - local.shape[0], *local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        local = torch.cat([local, pad], dim=0)
    out = torch.empty((world * max_n, *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local)
    if all(hi - lo == max_n for lo, hi in sizes):
        return out
    parts = [out[r * max_n: r * max_n + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    return torch.cat(parts, dim=0)


def data_parallel_step(step_fn: Callable[[Dict[str, torch.Tensor]], torch.Tensor],
                       inputs: Dict[str, torch.Tensor], global_batch: int,
                       rank: int, world: int) -> torch.Tensor:
    """Run `step_fn` on this rank's shard and return the gathered [global_batch, ...] latents."""
    local_out = step_fn(shard_inputs(inputs, global_batch, world, rank))
    return gather_latents(local_out, global_batch, world)
