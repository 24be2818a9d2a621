"""Host-side logic of the data-parallel denoise launcher, on the CPU: prompt sharding and the final
latent all-gather over gloo with world_size 2 (the N > 1 path of bench.py / pipeline.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stabletriton_b200.pipeline import gather_latents, shard_prompts


def test_shard_prompts_partitions_exactly():
    for total in (1, 2, 7, 8, 16, 33):
        for world in (1, 2, 3, 4, 8):
            shards = [shard_prompts(total, world, r) for r in range(world)]
            assert shards[0][0] == 0 and shards[-1][1] == total
            for (a, b), (c, d) in zip(shards, shards[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in shards]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == total


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_prompts(total, world, rank)
        # each rank "denoises" its prompts: latent p is filled with the value p
        local = torch.stack([torch.full((4, 8, 8), float(p)) for p in range(lo, hi)]) if hi > lo \
            else torch.zeros((0, 4, 8, 8))
        full = gather_latents(local, total)
        torch.save(full, os.path.join(out_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 5])
def test_gather_latents_gloo_world2(tmp_path, total):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    expect = torch.stack([torch.full((4, 8, 8), float(p)) for p in range(total)])
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        assert got.shape == expect.shape and torch.equal(got, expect)


def test_gather_latents_single_process_is_identity():
    x = torch.randn(3, 4, 8, 8)
    assert gather_latents(x, 3) is x
