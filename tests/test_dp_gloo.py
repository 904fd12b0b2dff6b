"""The N > 1 host logic (batch sharding + final latent all-gather) on CPU: world_size 2, gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mixdq_b200 import dp


def test_shard_bounds_cover_batch():
    for gb in (1, 2, 5, 8, 64):
        for world in (1, 2, 4, 8):
            spans = [dp.shard_bounds(gb, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == gb
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, gb, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = dp.init_distributed("gloo")
    assert (r, w) == (rank, world)
    g = torch.Generator().manual_seed(0)
    inputs = {"sample": torch.randn(gb, 4, 8, 8, generator=g),
              "encoder_hidden_states": torch.randn(gb, 7, 16, generator=g),
              "timestep": torch.tensor(999.0)}

    def step(local):
        assert local["timestep"].dim() == 0
        lo, hi = dp.shard_bounds(gb, world, rank)
        assert local["sample"].shape[0] == hi - lo
        return local["sample"] * 2 + local["encoder_hidden_states"].sum(dim=(1, 2))[:, None, None, None]

    out = dp.data_parallel_step(step, inputs, gb, rank, world)
    want = inputs["sample"] * 2 + inputs["encoder_hidden_states"].sum(dim=(1, 2))[:, None, None, None]
    q.put((rank, bool(torch.equal(out, want)), tuple(out.shape)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("gb", [4, 5])
def test_data_parallel_step_world2_gloo(gb):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, gb, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res) and all(r[2] == (gb, 4, 8, 8) for r in res)
