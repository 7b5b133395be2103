"""Pipelined mode on N GPUs (BASELINE config #5): mixer -> polyphase FIR bank (256 x 12) -> 256-point FFT -> gain, one
group of consecutive blocks per GPU. An edge that crosses GPUs is a buffer in the consumer's HBM that the producer's last
kernel stores into over NVLink (multigpu.PeerStoreChain, the default), or an NCCL send/recv pair per chunk (--transport nccl).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/bench_pipeline.py [--chunks K] [--chunk-samples S]

N = 1 runs the whole chain on one GPU (the baseline the pipeline is compared with), N = 2 splits it in two stages, N = 4 in
four, N = 8 runs two four-stage pipelines side by side. The last rank of every pipeline re-runs the first chunks of the
whole chain locally and compares bit for bit. Prints one JSON line (rank 0)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import gnuradio4_b200 as gr4
from gnuradio4_b200 import multigpu

M, P = 256, 12


def make_blocks(device_index):
    dom = f"gpu:cuda:{device_index}"
    proto = gr4.fir_generate(M * P, "Kaiser", 1.0 / (2 * M), beta=8.0)
    rot = gr4.Rotator(phase_increment=0.6283185, compute_domain=dom)
    chan = gr4.PolyphaseChannelizer(proto, M, compute_domain=dom)
    gain = gr4.MultiplyConst(value=0.5 + 0.25j, compute_domain=dom)
    return [lambda x, out: rot.process_bulk(x, out=out), lambda x, out: chan.filter_stage(x, out=out), lambda x, out: chan.fft_stage(x, out=out), lambda x, out: gain.process_bulk(x, out=out)]


def run_pipeline(rank, world, local, device, chunks=32, chunk_samples=1 << 24, verify_chunks=2, transport="peer"):
    """The measurement itself, on an initialised process group (bench.py calls it for its `workloads.pipeline` entry).
    Returns the result dict on rank 0, None elsewhere."""
    if world not in (1, 2, 4, 8):  # whole pipelines only: 1, 2 or 4 stages, two 4-stage pipelines side by side at 8
        return None
    n = chunk_samples // M * M
    n_stages = 1 if world == 1 else (2 if world == 2 else 4)
    active, group = world, None
    per_stage = 4 // n_stages
    blocks = make_blocks(local)
    pipeline, stage = multigpu.stage_assignment(n_stages, active)[rank]
    mine = blocks[stage * per_stage : (stage + 1) * per_stage]
    scratch = [[torch.empty(n, dtype=torch.complex64, device=device) for _ in mine] for _ in range(2)]  # two-deep outputs per block

    def stage_fn(x, k, out=None):  # the stage's last block writes into `out` (an edge slot on the next GPU) when given
        for b, fn in enumerate(mine):
            x = fn(x, out if out is not None and b == len(mine) - 1 else scratch[k % 2][b])
        return x

    gen = torch.Generator(device=device)
    gen.manual_seed(1234 + pipeline)
    src = torch.empty(2 * n, dtype=torch.complex64, device=device)  # the source replays two resident chunks (inputs >> L2)
    torch.view_as_real(src).uniform_(-1, 1, generator=gen)
    kept = {}

    def source(k):
        return src[(k % 2) * n : (k % 2 + 1) * n]

    def sink(k, y):
        if k < verify_chunks:
            kept[k] = y.clone()

    if transport == "peer":
        chain = multigpu.PeerStoreChain([stage_fn] * n_stages, in_shapes=[(n,)] * n_stages, dtype=torch.complex64, device=device, world=active)
    else:
        chain = multigpu.PipelinedChain([lambda x, k: stage_fn(x, k)] * n_stages, in_shapes=[(n,)] * n_stages, dtype=torch.complex64, device=device, world=active)
    chain.run(2, source=source, sink=sink)  # warm-up (also creates the NCCL channels); block state carries on
    kept.clear()
    torch.cuda.synchronize()
    dist.barrier(group=group)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    chain.run(chunks, source=source, sink=sink)
    stop.record()
    torch.cuda.synchronize()
    dist.barrier(group=group)
    t = torch.tensor([start.elapsed_time(stop)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    ms = float(t.item())

    # bit-for-bit check on the last rank of each pipeline: whole chain locally, same state history (2 warm-up chunks first)
    ok = True
    if chain.next is None and n_stages > 1:
        tmp = [torch.empty(n, dtype=torch.complex64, device=device) for _ in range(4)]
        # replay precisely: warm-up consumed chunks (0, 1); the timed run starts again at chunk index 0
        local_blocks = make_blocks(local)
        seq = [0, 1] + list(range(verify_chunks))
        for i, k in enumerate(seq):
            x = source(k)
            for b, fn in enumerate(local_blocks):
                x = fn(x, tmp[b])
            if i >= 2:
                ok = ok and torch.equal(torch.view_as_real(x).view(torch.int32), torch.view_as_real(kept[k]).view(torch.int32))
    flag = torch.tensor([1 if ok else 0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    pipelines = active // n_stages
    if transport == "peer":
        chain.close()
    if rank != 0:
        return None
    samples = chunks * n * pipelines
    return {"workload": "pfb256x12_channelizer_pipeline", "n_gpus": world, "gpus_used": active, "stages": n_stages, "pipelines": pipelines, "chunks": chunks, "chunk_samples": n, "edge_transport": chain.transport,
            "ms": ms, "GS/s": samples / ms / 1e6, "edge_GB/s_per_edge": 8.0 * chunks * n / ms / 1e6 if n_stages > 1 else 0.0, "bit_identical_to_single_gpu_chain": bool(flag.item())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chunks", type=int, default=32)
    ap.add_argument("--chunk-samples", type=int, default=1 << 24)
    ap.add_argument("--verify-chunks", type=int, default=2)
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"], help="peer: the producer's kernel stores into the consumer's HBM (CUDA IPC); nccl: send/recv per chunk")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    gr4.load()
    result = run_pipeline(rank, world, local, device, args.chunks, args.chunk_samples, args.verify_chunks, args.transport)
    if rank == 0:
        print(json.dumps(result))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
