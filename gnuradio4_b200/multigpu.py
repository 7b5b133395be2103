"""Multi-GPU plumbing (one process per GPU, torch.distributed): channel sharding for independent flowgraph replicas and
the inter-GPU edge of the pipelined mode. NCCL on the GPUs, gloo in the CPU tests -- the code path is the same.

The reference has no multi-process mode; its analogue of the pipelined mode is the multiThreaded scheduler policy that
deals blocks out to worker threads with CircularBuffers as the hand-off queues (core/include/gnuradio-4.0/Scheduler.hpp:
1944-1951). Here the hand-off between two GPUs is one send/recv pair per chunk over NVLink.
"""
import torch
import torch.distributed as dist


def channel_assignment(n_channels, world_size):
    """Independent channels share nothing: channel c runs on rank c mod world_size."""
    return [[c for c in range(n_channels) if c % world_size == r] for r in range(world_size)]


def stage_assignment(n_stages, world_size):
    """Pipelined mode: consecutive blocks on consecutive GPUs; with more GPUs than stages, several pipelines run side by
    side. Returns for every rank (pipeline index, stage index)."""
    if world_size % n_stages != 0:
        raise ValueError("world size must be a multiple of the number of pipeline stages")
    return [(r // n_stages, r % n_stages) for r in range(world_size)]


def max_over_ranks(value, device=None):
    """Timing rule of the bench contract: the slowest rank defines the step."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class PipelineEdge:
    """One edge of the flowgraph that crosses GPUs: the producer rank sends each published chunk, the consumer rank
    receives it into its own ring span. ncclSend / ncclRecv underneath (gloo send / recv on CPU)."""

    def __init__(self, producer_rank, consumer_rank, group=None):
        self.producer_rank, self.consumer_rank, self.group = producer_rank, consumer_rank, group
        self.rank = dist.get_rank()
        self.chunks = 0
        self.bytes = 0

    def publish(self, chunk):
        assert self.rank == self.producer_rank
        dist.send(chunk, dst=self.consumer_rank, group=self.group)
        self.chunks += 1
        self.bytes += chunk.numel() * chunk.element_size()

    def get(self, out):
        assert self.rank == self.consumer_rank
        dist.recv(out, src=self.producer_rank, group=self.group)
        self.chunks += 1
        self.bytes += out.numel() * out.element_size()
        return out
