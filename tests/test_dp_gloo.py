"""Data-parallel wrapper on CPU: world_size 2 over gloo.  Points are sharded, cells and head
replicated, one flat all-reduce of the gradients; the result must equal the single-process
step over all points.  The sampler here is the CPU oracle (tests may use it); on the GPU
box the same wrapper drives the CUDA op over NCCL (bench.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cosinesampler_b200 import chain, dp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem():
    gen = torch.Generator().manual_seed(11)
    cells = torch.rand(2, 4, 8, 8, generator=gen, dtype=torch.float64)
    coords = torch.rand(37, 2, generator=gen, dtype=torch.float64) * 1.9 - 0.95
    return cells, coords


def _sampler(c, g):
    from oracle.grid_sampler_oracle import grid_sample_2d
    return grid_sample_2d(c, g, step="cosine", offset=True)


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cells0, coords = _problem()
        cells = torch.nn.Parameter(cells0.clone())
        head = chain.make_head(4, seed=3, dtype=torch.float64)
        P = coords.shape[0]
        s, e = dp.shard_range(P, rank, world)
        step = dp.PointShardedStep(_sampler, cells, head, residual="helmholtz", chunk=10)
        step.zero_grad()
        loss = step.step([coords[s:e, 0:1], coords[s:e, 1:2]], P)
        total = loss.clone()
        dist.all_reduce(total)
        if rank == 0:
            torch.save({"cells": cells.grad, "head": [p.grad for p in head.parameters()],
                        "loss": total}, out_path)
    finally:
        dist.destroy_process_group()


def test_point_sharded_step_matches_single_process(tmp_path):
    world = 2
    out_path = str(tmp_path / "dp.pt")
    mp.spawn(_worker, args=(world, _free_port(), out_path), nprocs=world, join=True)
    got = torch.load(out_path)

    cells0, coords = _problem()
    cells = torch.nn.Parameter(cells0.clone())
    head = chain.make_head(4, seed=3, dtype=torch.float64)
    loss = chain.training_step(_sampler, cells, [coords[:, 0:1], coords[:, 1:2]], head,
                               residual="helmholtz")
    torch.testing.assert_close(got["loss"], loss, rtol=1e-10, atol=1e-12)
    torch.testing.assert_close(got["cells"], cells.grad, rtol=1e-9, atol=1e-12)
    for g, p in zip(got["head"], head.parameters()):
        torch.testing.assert_close(g, p.grad, rtol=1e-9, atol=1e-12)


def test_allreduce_grads_is_a_noop_without_a_group():
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    dp.allreduce_grads([p])
    assert torch.equal(p.grad, torch.full((3,), 2.0))
