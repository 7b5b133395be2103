"""The inter-process edge of the pipelined mode (multigpu.PeerStoreChain): two PROCESSES, consecutive blocks of one chain,
the producer's kernel storing into the consumer's memory through a CUDA IPC mapping, cursors as stream-ordered counters.
Runs on one GPU (both processes on cuda:0 -- IPC does not care that the "peer" is the same device), so the protocol is
covered by the single-GPU `pytest -m gpu` run; scripts/bench_pipeline.py runs it across GPUs over NVLink."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, chunks, result_path):
    import torch.distributed as dist

    import gnuradio4_b200 as gr4
    from gnuradio4_b200 import multigpu

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    device = torch.device("cuda", 0)
    taps = gr4.fir_generate(127, "Hamming", 0.1)
    blocks = [gr4.MultiplyConst(value=0.5 - 0.25j), gr4.fir_filter(b=taps), gr4.Rotator(phase_increment=0.3)]
    scratch = [torch.empty(n, dtype=torch.complex64, device=device) for _ in range(2)]

    def stage(x, k, out):
        return blocks[rank].process_bulk(x, out=out if out is not None else scratch[k % 2])

    gen = torch.Generator(device=device)
    gen.manual_seed(7)
    src = torch.empty(chunks * n, dtype=torch.complex64, device=device)
    torch.view_as_real(src).uniform_(-1, 1, generator=gen)
    got = []
    chain = multigpu.PeerStoreChain([stage] * world, in_shapes=[(n,)] * world, dtype=torch.complex64, device=device)
    half = chunks // 2  # two runs: the counters carry on
    for first, count in ((0, half), (half, chunks - half)):
        chain.run(count, source=lambda i, f=first: src[(f + i) * n : (f + i + 1) * n], sink=lambda i, y: got.append(y.clone()))
    torch.cuda.synchronize()
    if rank == world - 1:
        whole = [gr4.MultiplyConst(value=0.5 - 0.25j), gr4.fir_filter(b=taps), gr4.Rotator(phase_increment=0.3)][:world]
        want = []
        for c in range(chunks):
            x = src[c * n : (c + 1) * n]
            for b in whole:
                x = b.process_bulk(x)
            want.append(x)
        same = all(torch.equal(torch.view_as_real(a).view(torch.int32), torch.view_as_real(b).view(torch.int32)) for a, b in zip(got, want))
        with open(result_path, "w") as f:
            f.write(f"{len(got)} {int(same)}")
    chain.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_store_chain_two_and_three_processes(tmp_path, world):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import torch.multiprocessing as mp

    result = tmp_path / "result.txt"
    chunks, n = 9, 1 << 16
    mp.spawn(_worker, args=(world, _free_port(), n, chunks, str(result)), nprocs=world, join=True)
    count, same = result.read_text().split()
    assert int(count) == chunks and same == "1", "the pipelined chain differs from the chain in one process"
