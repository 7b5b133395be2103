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


class PipelinedChain:
    """Pipelined mode (BASELINE config #5): consecutive blocks of one linear chain live on consecutive ranks
    (GPUs); every edge that crosses ranks carries one chunk per step as a send/recv pair (ncclSend / ncclRecv over NVLink;
    gloo on CPU). With world = k * n_stages, k pipelines run side by side on independent channels.

    `stages[i]` is a callable `(chunk_tensor, chunk_index) -> tensor` (a block's process_bulk; it may write into output
    buffers it reuses every second chunk, the chain waits for the send of chunk k-2 before calling it); stage 0 is fed by
    `source(chunk_index) -> tensor`, the last stage hands its result to `sink(chunk_index, tensor)`.
    Receives are posted one chunk ahead into a two-deep buffer, sends are asynchronous: the transfer of chunk k+1 overlaps
    the work on chunk k, as the producer/consumer threads of the reference's multi-threaded scheduler overlap through a
    CircularBuffer (Scheduler.hpp:1944-1951). Stream order carries every dependency; the host never blocks on a chunk."""

    transport = "nccl send/recv (one communicator per edge)"

    def __init__(self, stages, in_shapes, dtype, device, group=None, edge_groups=True, world=None):
        self.rank, self.world = dist.get_rank(), (world if world is not None else dist.get_world_size())  # world: the ranks that take part
        self.n_stages = len(stages)
        self.pipeline, self.stage = stage_assignment(self.n_stages, self.world)[self.rank]
        self.fn = stages[self.stage]
        self.prev = self.rank - 1 if self.stage > 0 else None
        self.next = self.rank + 1 if self.stage + 1 < self.n_stages else None
        # One communicator per edge: a middle stage receives chunk k+1 and sends chunk k at the same time, which needs
        # the two transfers on different NCCL streams (one process group = one stream per device). Every rank takes part
        # in the creation of every group, in the same order.
        self.recv_group = self.send_group = group
        if edge_groups and group is None and self.n_stages > 2:
            for a in range(self.world - 1):
                if (a + 1) % self.n_stages == 0:
                    continue  # no edge between the last stage of one pipeline and the first of the next
                g = dist.new_group([a, a + 1])
                if a + 1 == self.rank:
                    self.recv_group = g
                if a == self.rank:
                    self.send_group = g
        # input buffers of this stage (what the upstream rank sends): two deep
        self.inbox = [torch.empty(in_shapes[self.stage], dtype=dtype, device=device) for _ in range(2)] if self.prev is not None else None
        self.sent_bytes = 0
        self.received_bytes = 0

    def run(self, n_chunks, source=None, sink=None):
        pending_send = [None, None]
        keep_alive = [None, None]
        recv_req = None
        if self.prev is not None and n_chunks > 0:
            recv_req = dist.irecv(self.inbox[0], src=self.prev, group=self.recv_group)
        for k in range(n_chunks):
            if self.prev is None:
                chunk = source(k)
            else:
                recv_req.wait()
                chunk = self.inbox[k % 2]
                self.received_bytes += chunk.numel() * chunk.element_size()
                if k + 1 < n_chunks:  # the other half of the inbox was consumed by chunk k-1's work, already enqueued
                    recv_req = dist.irecv(self.inbox[(k + 1) % 2], src=self.prev, group=self.recv_group)
            if pending_send[k % 2] is not None:
                pending_send[k % 2].wait()  # a stage may reuse its output buffers with period 2: chunk k-2 must have left
                pending_send[k % 2] = None
            out = self.fn(chunk, k)
            if self.next is not None:
                keep_alive[k % 2] = out
                pending_send[k % 2] = dist.isend(out, dst=self.next, group=self.send_group)
                self.sent_bytes += out.numel() * out.element_size()
            elif sink is not None:
                sink(k, out)
        for req in pending_send:
            if req is not None:
                req.wait()


def bind_to_device_numa_node(device_index):
    """Pins the calling process to the CPU cores that sit next to `device_index` (the PCI device's `local_cpulist`), so
    that the pinned host buffers it allocates afterwards are first-touched on that NUMA node and host<->device copies do
    not cross the socket interconnect. One process per GPU: without it eight ranks share whichever node they started
    on. Returns the CPU set, or None when the topology cannot be read (the process is left as it was)."""
    import os

    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        domain = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{domain:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        with open(path) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                lo, hi = part.split("-")
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # no sysfs entry, no permission, not Linux: keep the default placement
        return None


# ---- edges between processes without a collective: the producer's kernel stores into the consumer's memory -------------------
def tensor_at(pointer, n_items, dtype, device):
    """A torch tensor over device memory this process did not allocate through torch (an IPC-mapped edge buffer of the next
    GPU): zero-copy through the CUDA array interface. Blocks then take it as `out=` like any other tensor."""
    words = {torch.complex64: 2, torch.float32: 1}[dtype] * int(n_items)

    class _Memory:
        __cuda_array_interface__ = {"shape": (words,), "typestr": "<f4", "data": (int(pointer), False), "version": 2, "strides": None}

    flat = torch.as_tensor(_Memory(), device=device)
    return torch.view_as_complex(flat.view(-1, 2)) if dtype == torch.complex64 else flat


class PeerStoreChain:
    """Pipelined mode with the edge INSIDE the producing kernel: consecutive blocks of one linear chain live on consecutive
    ranks (GPUs) like PipelinedChain, but an edge that crosses ranks is a two-chunk buffer in the CONSUMER's HBM which the
    producer maps through CUDA IPC; the last block of the producer's stage writes its output straight into it over NVLink
    (plain stores from the kernel: transfer and compute are one launch, no staging buffer, no send/recv). The edge's two
    cursors are 32-bit counters in device memory moved and awaited by the streams themselves (gr4b200_stream_write_value32
    / _wait_value32): `ready` in the consumer's memory counts the chunks that have landed, `free` in the producer's memory
    the chunks the consumer is done with. Neither host ever blocks or exchanges a message once the handles are swapped.

    `stages[i]` is a callable `(chunk_tensor, chunk_index, out_tensor) -> tensor`: it must leave the stage's result in
    `out_tensor` (an edge slot on the next GPU, or None on the last stage: then it returns its own buffer)."""

    transport = "peer store: the producer's kernel writes the consumer's HBM over NVLink (CUDA IPC mapping, stream-ordered counters)"

    def __init__(self, stages, in_shapes, dtype, device, world=None):
        import ctypes as C

        from . import _lib

        self._C, self._lib = C, _lib.load()
        self.rank, self.world = dist.get_rank(), (world if world is not None else dist.get_world_size())
        self.n_stages = len(stages)
        self.pipeline, self.stage = stage_assignment(self.n_stages, self.world)[self.rank]
        self.fn = stages[self.stage]
        self.prev = self.rank - 1 if self.stage > 0 else None
        self.next = self.rank + 1 if self.stage + 1 < self.n_stages else None
        self.device, self.dtype = device, dtype
        self.done = 0  # chunks since construction (the counters are absolute)
        item = {torch.complex64: 8, torch.float32: 4}[dtype]
        header = 256
        # what this rank owns: its inbox (ready counter + two slots) if it has a producer, its free counter if it has a consumer
        self._inbox = self._free = None
        mine = {"inbox": None, "free": None}
        if self.prev is not None:
            self.n_in = int(torch.Size(in_shapes[self.stage]).numel())
            self._inbox = _lib.check_ptr(self._lib.gr4b200_malloc(header + 2 * self.n_in * item), "malloc(edge inbox)")
            _lib.check(self._lib.gr4b200_memset(C.c_void_p(self._inbox), 0, header, None), "memset")
            mine["inbox"] = self._export(self._inbox)
            self.slots = [tensor_at(self._inbox + header + s * self.n_in * item, self.n_in, dtype, device) for s in range(2)]
        if self.next is not None:
            self._free = _lib.check_ptr(self._lib.gr4b200_malloc(header), "malloc(edge counter)")
            _lib.check(self._lib.gr4b200_memset(C.c_void_p(self._free), 0, header, None), "memset")
            mine["free"] = self._export(self._free)
        torch.cuda.synchronize()
        everyone = [None] * dist.get_world_size()
        dist.all_gather_object(everyone, mine)
        self._remote_inbox = self._remote_free = None
        if self.next is not None:  # map the consumer's inbox: its ready counter and the two slots we store into
            n_out = int(torch.Size(in_shapes[self.stage + 1]).numel())
            self._remote_inbox = self._open(everyone[self.next]["inbox"])
            self.remote_slots = [tensor_at(self._remote_inbox + header + s * n_out * item, n_out, dtype, device) for s in range(2)]
        if self.prev is not None:  # map the producer's free counter
            self._remote_free = self._open(everyone[self.prev]["free"])
        self.sent_bytes = self.received_bytes = 0

    def _export(self, pointer):
        handle = (self._C.c_ubyte * 64)()
        from . import _lib

        _lib.check(self._lib.gr4b200_ipc_export(self._C.c_void_p(pointer), handle), "ipc_export")
        return bytes(handle)

    def _open(self, handle):
        from . import _lib

        buf = (self._C.c_ubyte * 64).from_buffer_copy(handle)
        return _lib.check_ptr(self._lib.gr4b200_ipc_open(buf), "ipc_open")

    def run(self, n_chunks, source=None, sink=None):
        from . import _lib

        C, lib = self._C, self._lib
        for i in range(n_chunks):
            k = self.done + i
            stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            if self.prev is None:
                chunk = source(i)
            else:  # chunk k has landed in slot k % 2 once the producer's counter reads k + 1
                _lib.check(lib.gr4b200_stream_wait_value32(stream, C.c_void_p(self._inbox), k + 1), "wait(ready)")
                chunk = self.slots[k % 2]
                self.received_bytes += chunk.numel() * chunk.element_size()
            out = None
            if self.next is not None:
                if k >= 2:  # slot k % 2 still holds chunk k - 2 until the consumer's counter reads k - 1
                    _lib.check(lib.gr4b200_stream_wait_value32(stream, C.c_void_p(self._free), k - 1), "wait(free)")
                out = self.remote_slots[k % 2]
            result = self.fn(chunk, i, out)
            if self.next is not None:
                _lib.check(lib.gr4b200_stream_write_value32(stream, C.c_void_p(self._remote_inbox), k + 1), "write(ready)")
                self.sent_bytes += out.numel() * out.element_size()
            elif sink is not None:
                sink(i, result)
            if self.prev is not None:  # our work on chunk k is enqueued: its slot is free once the stream gets here
                _lib.check(lib.gr4b200_stream_write_value32(stream, C.c_void_p(self._remote_free), k + 1), "write(free)")
        self.done += n_chunks

    def close(self):
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier()  # nobody unmaps or frees while a neighbour's stream may still touch the memory
        for mapped in (self._remote_inbox, self._remote_free):
            if mapped:
                self._lib.gr4b200_ipc_close(self._C.c_void_p(mapped))
        for owned in (self._inbox, self._free):
            if owned:
                self._lib.gr4b200_free(self._C.c_void_p(owned))
        self._remote_inbox = self._remote_free = self._inbox = self._free = None
