"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: channel sharding, the max-over-ranks timing rule and the
inter-GPU pipeline edge (NCCL on the GPUs, gloo here; same code)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gnuradio4_b200 import multigpu


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # replicas: every rank owns its channels, no exchange; the step time is the slowest rank's
        mine = multigpu.channel_assignment(8, world)[rank]
        slowest = multigpu.max_over_ranks(10.0 + rank)
        # pipelined: stage 0 (rank 0) multiplies by 2 and ships each chunk to stage 1 (rank 1), which adds 1
        pipeline, stage = multigpu.stage_assignment(2, world)[rank]
        edge = multigpu.PipelineEdge(0, 1)
        n_chunks, chunk = 5, 1024
        rng = np.random.default_rng(7)
        x = torch.from_numpy((rng.uniform(-1, 1, n_chunks * chunk) + 1j * rng.uniform(-1, 1, n_chunks * chunk)).astype(np.complex64))
        out = []
        for k in range(n_chunks):
            if stage == 0:
                edge.publish(torch.view_as_real(x[k * chunk : (k + 1) * chunk] * 2).contiguous())
            else:
                buf = torch.empty((chunk, 2), dtype=torch.float32)
                out.append(torch.view_as_complex(edge.get(buf)) + 1)
        ok = True
        if stage == 1:
            ok = torch.equal(torch.cat(out), x * 2 + 1)
        results[rank] = (mine, slowest, pipeline, stage, edge.chunks, edge.bytes, ok)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_pipeline_edge():
    world = 2
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(worker, args=(world, free_port(), results), nprocs=world, join=True)
    assert results[0][0] == [0, 2, 4, 6] and results[1][0] == [1, 3, 5, 7]
    assert results[0][1] == 11.0 and results[1][1] == 11.0  # max over ranks on both
    assert results[0][2:4] == (0, 0) and results[1][2:4] == (0, 1)
    assert results[0][4] == 5 and results[1][4] == 5 and results[0][5] == results[1][5] == 5 * 1024 * 8
    assert results[1][6]


def test_assignments():
    assert multigpu.channel_assignment(8, 8) == [[c] for c in range(8)]
    assert multigpu.channel_assignment(3, 2) == [[0, 2], [1]]
    assert multigpu.stage_assignment(4, 8) == [(0, 0), (0, 1), (0, 2), (0, 3), (1, 0), (1, 1), (1, 2), (1, 3)]
    with pytest.raises(ValueError):
        multigpu.stage_assignment(3, 8)


def chain_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_chunks, chunk = 7, 512
        outs = [torch.empty(chunk), torch.empty(chunk)]  # stage 0 reuses its output buffers with period 2, like a device stage

        def first(x, k):
            torch.mul(x, 2.0, out=outs[k % 2])
            return outs[k % 2]

        stages = [first, lambda x, k: x + float(k)]  # the second stage depends on the chunk index: order matters
        chain = multigpu.PipelinedChain(stages, in_shapes=[(chunk,), (chunk,)], dtype=torch.float32, device="cpu")
        pipeline = chain.pipeline
        gen = torch.Generator().manual_seed(100 + pipeline)
        data = torch.rand(n_chunks * chunk, generator=gen)  # both ranks of a pipeline can regenerate the source
        got = []
        chain.run(n_chunks, source=lambda k: data[k * chunk : (k + 1) * chunk].clone(), sink=lambda k, y: got.append(y.clone()))
        ok = True
        if chain.next is None:
            want = torch.cat([data[k * chunk : (k + 1) * chunk] * 2.0 + float(k) for k in range(n_chunks)])
            ok = torch.equal(torch.cat(got), want)
        results[rank] = (chain.pipeline, chain.stage, chain.sent_bytes, chain.received_bytes, ok)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_pipelined_chain_over_ranks(world):
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(chain_worker, args=(world, free_port(), results), nprocs=world, join=True)
    for rank in range(world):
        pipeline, stage, sent, received, ok = results[rank]
        assert (pipeline, stage) == (rank // 2, rank % 2)
        assert ok
        assert (sent, received) == ((7 * 512 * 4, 0) if stage == 0 else (0, 7 * 512 * 4))


def four_stage_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_chunks, chunk = 6, 256
        stages = [lambda x, k: x + 1.0, lambda x, k: x * 3.0, lambda x, k: x - float(k), lambda x, k: x * 0.5]
        chain = multigpu.PipelinedChain(stages, in_shapes=[(chunk,)] * 4, dtype=torch.float32, device="cpu")  # one group per edge
        data = torch.arange(n_chunks * chunk, dtype=torch.float32)
        got = []
        chain.run(n_chunks, source=lambda k: data[k * chunk : (k + 1) * chunk].clone(), sink=lambda k, y: got.append(y.clone()))
        ok = True
        if chain.next is None:
            want = torch.cat([((data[k * chunk : (k + 1) * chunk] + 1.0) * 3.0 - float(k)) * 0.5 for k in range(n_chunks)])
            ok = torch.equal(torch.cat(got), want)
        results[rank] = (chain.stage, ok, chain.recv_group is not chain.send_group or chain.stage in (0, 3))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_four_stage_pipeline_with_one_group_per_edge():
    manager = mp.Manager()
    results = manager.dict()
    mp.spawn(four_stage_worker, args=(4, free_port(), results), nprocs=4, join=True)
    assert [results[r][0] for r in range(4)] == [0, 1, 2, 3]
    assert all(results[r][1] and results[r][2] for r in range(4))
