"""gr::Graph / gr::scheduler::Simple shaped driver for device flowgraphs (linear chains on this path).

Mirrors the calls a reference user makes (core/include/gnuradio-4.0/Graph.hpp:425 emplaceBlock, :596 connect;
core/include/gnuradio-4.0/Scheduler.hpp:394 exchange, :581 runAndWait) with a different engine underneath:
  * every edge is an HBM ring (gr4b200_ring_*, the device replacement of CircularBuffer<T>), default 2 chunks deep;
  * a block's work chunk is one asynchronous C-ABI launch on the compute stream (Block::workInternal ->
    dispatchProcessing, Block.hpp:2028-2172, with stream order instead of host-thread order);
  * host sources / sinks are pinned buffers; H2D, compute and D2H run on three streams and overlap chunk by chunk, the
    ring's publish/consume events carry the dependencies (the reference's "explicit conversion blocks" between domains).
Chunk sizes obey the reference rule for Resampling blocks: whole multiples of input_chunk_size (Block.hpp:1610-1635).
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import Gr4b200Error, check, check_ptr


class HostBuffer:
    """Pinned host memory (gr4b200_malloc_host) exposed as a numpy array."""

    def __init__(self, n_items, dtype):
        self._lib = _lib.load()
        self.dtype = np.dtype(dtype)
        self.nbytes = int(n_items) * self.dtype.itemsize
        self.ptr = check_ptr(self._lib.gr4b200_malloc_host(max(self.nbytes, 1)), "malloc_host")
        self.array = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(self.ptr)).view(self.dtype)

    def close(self):
        if self.ptr:
            self.array = None
            self._lib.gr4b200_free_host(self.ptr)
            self.ptr = None

    def __del__(self):
        self.close()


class Graph:
    def __init__(self):
        self.blocks = []
        self.edges = []  # (src, dst, min_buffer_size in items or None)

    def emplaceBlock(self, block_type, **settings):  # noqa: N802 -- reference spelling
        block = block_type(**settings)
        block._grc_settings = dict(settings)  # what save_grc writes back
        self.blocks.append(block)
        return block

    def connect(self, src, dst, minBufferSize=None):  # noqa: N803
        if src not in self.blocks or dst not in self.blocks:
            raise Gr4b200Error("connect: both blocks must have been emplaced in this graph")
        if any(e[0] is src for e in self.edges) or any(e[1] is dst for e in self.edges):
            raise Gr4b200Error("connect: this path supports one edge per port (linear chains)")
        self.edges.append((src, dst, minBufferSize))
        return True

    def chain(self):
        """Blocks in stream order; raises if the graph is not a single linear chain."""
        if not self.blocks:
            return []
        heads = [b for b in self.blocks if not any(e[1] is b for e in self.edges)]
        if len(heads) != 1:
            raise Gr4b200Error("graph is not a single linear chain")
        order, current = [heads[0]], heads[0]
        while True:
            nxt = [e[1] for e in self.edges if e[0] is current]
            if not nxt:
                break
            current = nxt[0]
            order.append(current)
        if len(order) != len(self.blocks):
            raise Gr4b200Error("graph has unconnected blocks")
        return order


class Simple:
    """scheduler::Simple for a linear device chain fed from / drained to host memory."""

    def __init__(self, graph=None, chunk_items=1 << 22, device=None):
        self._lib = _lib.load()
        self.device = device  # None: the device of the graph's blocks
        self.chunk_items = int(chunk_items)
        self.graph = None
        self._rings = []
        self._rings_aligned = True
        self._streams = None
        self.launches = 0
        if graph is not None:
            self.exchange(graph)

    def exchange(self, graph):
        self.graph = graph
        self._chain = graph.chain()
        devices = {block.device for block in self._chain}
        if self.device is None and len(devices) == 1:
            self.device = devices.pop()
        elif devices != {self.device}:
            raise Gr4b200Error(f"scheduler on cuda:{self.device} but the chain's blocks live on {sorted('cuda:%d' % d for d in devices)}: one device per chain on this path")
        check(self._lib.gr4b200_init(self.device), "init")
        # the chunk must be a whole number of every block's input_chunk_size along the chain
        lcm, ratio_num, ratio_den = 1, 1, 1
        for block in self._chain:
            need = block.input_chunk_size * ratio_den // math.gcd(block.input_chunk_size * ratio_den, ratio_num)
            lcm = lcm * need // math.gcd(lcm, need)
            ratio_num *= block.output_chunk_size
            ratio_den *= block.input_chunk_size
        self.unit_items = lcm  # smallest input count every block of the chain can work on
        self.chunk_items = max(lcm, self.chunk_items // lcm * lcm)
        self.items_left_over = 0
        return graph

    def _ensure(self):
        check(self._lib.gr4b200_init(self.device), "init")  # streams, rings and launches belong to this scheduler's device
        if self._rings and not self._rings_aligned:  # a ragged last chunk left the cursors off the chunk grid: start over
            for ring in self._rings:
                self._lib.gr4b200_ring_destroy(ring)
            self._rings = []
        if self._streams is None:
            self._streams = [check_ptr(self._lib.gr4b200_stream_create(), "stream_create") for _ in range(3)]
        if not self._rings:
            n = self.chunk_items
            sizes = [(n, self._chain[0].in_item_bytes)]
            for block in self._chain:
                n = block.n_outputs_for(n)
                sizes.append((n, block.out_item_bytes))
            self._chunk_sizes = sizes
            # ring depth 2 chunks: the producer fills one while the consumer drains the other
            self._rings = [check_ptr(self._lib.gr4b200_ring_create(self.device, 2 * items * item_bytes, 0), "ring_create") for items, item_bytes in sizes]
            self._rings_aligned = True

    def n_outputs_for(self, n_in):
        n = n_in
        for block in self._chain:
            n = block.n_outputs_for(n)
        return n

    def out_item_bytes(self):
        return self._chain[-1].out_item_bytes

    def runAndWait(self, host_in, host_out):  # noqa: N802
        """Streams host_in (pinned numpy array of input items) through the chain into host_out (pinned); returns the bytes
        written. Work chunks are whole multiples of every block's input_chunk_size (Block.hpp:1610-1635): a remainder
        shorter than that unit is not consumed (`items_left_over`), as the reference leaves it in the edge buffer."""
        self._ensure()
        lib = self._lib
        h2d, compute, d2h = self._streams
        in_bytes = self._chain[0].in_item_bytes
        n_total = host_in.shape[0] // self.unit_items * self.unit_items
        self.items_left_over = host_in.shape[0] - n_total
        need_out = self.n_outputs_for(n_total) * self._chain[-1].out_item_bytes
        if host_out.nbytes < need_out:
            raise Gr4b200Error(f"runAndWait: the output buffer holds {host_out.nbytes} bytes, {n_total} input items produce {need_out}")
        if n_total % self.chunk_items != 0:
            self._rings_aligned = False
        src_ptr = host_in.ctypes.data
        dst_ptr = host_out.ctypes.data
        out_bytes_done = 0
        self.launches = 0
        for first in range(0, n_total, self.chunk_items):
            n = min(self.chunk_items, n_total - first)
            # source: host -> ring 0
            ring = self._rings[0]
            nbytes = n * in_bytes
            dev = lib.gr4b200_ring_reserve(ring, nbytes, h2d)
            while not dev:  # ring full: wait for the consumer side to drain (back-pressure)
                check(lib.gr4b200_stream_synchronize(compute), "sync")
                dev = check_ptr(lib.gr4b200_ring_reserve(ring, nbytes, h2d), "ring_reserve")
            check(lib.gr4b200_copy_h2d(dev, src_ptr + first * in_bytes, nbytes, h2d), "copy_h2d")
            check(lib.gr4b200_ring_publish(ring, nbytes, h2d), "ring_publish")
            # blocks
            for k, block in enumerate(self._chain):
                n_out = block.n_outputs_for(n)
                rin, rout = self._rings[k], self._rings[k + 1]
                in_nbytes, out_nbytes = n * block.in_item_bytes, n_out * block.out_item_bytes
                in_dev = check_ptr(lib.gr4b200_ring_get(rin, in_nbytes, compute), "ring_get")
                out_dev = lib.gr4b200_ring_reserve(rout, out_nbytes, compute)
                while not out_dev:
                    check(lib.gr4b200_stream_synchronize(d2h), "sync")
                    out_dev = check_ptr(lib.gr4b200_ring_reserve(rout, out_nbytes, compute), "ring_reserve")
                block.launch(compute, in_dev, out_dev, n)  # n > 0 and a multiple of the chain's unit: n_out > 0
                self.launches += 1
                check(lib.gr4b200_ring_publish(rout, out_nbytes, compute), "ring_publish")
                check(lib.gr4b200_ring_consume(rin, in_nbytes, compute), "ring_consume")
                n = n_out
            # sink: last ring -> host
            ring = self._rings[-1]
            nbytes = n * self._chain[-1].out_item_bytes
            dev = check_ptr(lib.gr4b200_ring_get(ring, nbytes, d2h), "ring_get")
            check(lib.gr4b200_copy_d2h(dst_ptr + out_bytes_done, dev, nbytes, d2h), "copy_d2h")
            check(lib.gr4b200_ring_consume(ring, nbytes, d2h), "ring_consume")
            out_bytes_done += nbytes
        for s in (h2d, compute, d2h):
            check(lib.gr4b200_stream_synchronize(s), "stream_synchronize")
        return out_bytes_done

    def close(self):
        for ring in self._rings:
            self._lib.gr4b200_ring_destroy(ring)
        self._rings = []
        if self._streams:
            for s in self._streams:
                self._lib.gr4b200_stream_destroy(s)
            self._streams = None


# ---- .grc files (core/include/gnuradio-4.0/Graph_yaml_importer.hpp:88-380: `blocks: [{id, parameters: {name, ...}}]`,
# ---- `connections: [[source name, port, destination name, port, (minBufferSize)]]`) ------------------------------------
def parse_grc(text):
    """The two lists of a .grc document as plain Python data: [(block type without template arguments, template arguments,
    name, settings)], [(source name, source port, destination name, destination port, min buffer size or None)].
    Value type tags of the reference's YAML dialect (`!!float32 43`) are read as the plain Python value."""
    import yaml

    class Loader(yaml.SafeLoader):
        pass

    def tagged(loader, suffix, node):
        if not isinstance(node, yaml.ScalarNode):
            return loader.construct_sequence(node) if isinstance(node, yaml.SequenceNode) else loader.construct_mapping(node)
        value = loader.construct_scalar(node)
        if suffix.startswith(("float", "complex")):
            return float(value)
        if suffix.startswith(("int", "uint")):
            return int(value)
        return value

    Loader.add_multi_constructor("tag:yaml.org,2002:", tagged)
    doc = yaml.load(text, Loader=Loader) or {}
    blocks, connections = [], []
    for entry in doc.get("blocks") or []:
        if "id" not in entry:
            raise Gr4b200Error("grc: block without 'id'")
        full = str(entry["id"]).strip()
        base, _, args = full.partition("<")
        settings = dict(entry.get("parameters") or {})
        name = settings.pop("name", None)
        if name is None:
            raise Gr4b200Error(f"grc: block '{full}' has no parameters.name")  # Graph_yaml_importer.hpp:102
        blocks.append((base.strip(), args.rstrip("> ").strip(), str(name), settings))
    for conn in doc.get("connections") or []:
        if len(conn) < 4:
            raise Gr4b200Error(f"grc: unable to parse connection ({len(conn)} instead of >= 4 elements)")  # :293-296
        connections.append((str(conn[0]), conn[1], str(conn[2]), conn[3], int(conn[4]) if len(conn) > 4 else None))
    return blocks, connections


def _grc_registry():
    from . import blocks as b

    return {"gr::filter::fir_filter": b.fir_filter, "gr::filter::BasicDecimatingFilter": b.BasicDecimatingFilter, "gr::filter::Decimator": b.Decimator, "gr::blocks::fft::FFT": b.FFT,
            "gr::blocks::math::AddConst": b.AddConst, "gr::blocks::math::SubtractConst": b.SubtractConst, "gr::blocks::math::MultiplyConst": b.MultiplyConst, "gr::blocks::math::DivideConst": b.DivideConst,
            "gr::blocks::math::Rotator": b.Rotator}


def load_grc(text, compute_domain="gpu:cuda", registry=None):
    """gr::loadGrc for the blocks of this path: builds a Graph from a .grc document. Blocks without a `compute_domain`
    parameter get `compute_domain` (this package has no host path); unknown block types, element types other than
    complex<float> / float and unknown block names in a connection are errors, as in the reference's importer."""
    registry = registry or _grc_registry()
    specs, connections = parse_grc(text)
    graph, by_name = Graph(), {}
    for base, args, name, settings in specs:
        if base not in registry:
            raise Gr4b200Error(f"grc: block type '{base}' is not provided by this package")
        element = args.split(",")[0].strip()
        if element not in ("", "complex64", "std::complex<float", "std::complex<float>", "float32", "float"):
            raise Gr4b200Error(f"grc: '{base}<{args}>': this path runs complex<float> and float streams")
        settings.setdefault("compute_domain", compute_domain)
        if name in by_name:
            raise Gr4b200Error(f"grc: duplicate block name '{name}'")
        by_name[name] = graph.emplaceBlock(registry[base], **settings)
        by_name[name].name = name
    for src, src_port, dst, dst_port, min_buffer in connections:
        for block_name in (src, dst):
            if block_name not in by_name:
                raise Gr4b200Error(f"grc: unknown block '{block_name}'")  # :311-313
        for port, names in ((src_port, (0, "out")), (dst_port, (0, "in"))):
            if port not in names:
                raise Gr4b200Error(f"grc: port {port!r}: the blocks of this path have one input 'in' and one output 'out'")
        graph.connect(by_name[src], by_name[dst], minBufferSize=min_buffer)
    return graph


def save_grc(graph):
    """gr::saveGrc: the graph back as a .grc document (block type, name and the settings the block was built with)."""
    import yaml

    names = {v: k for k, v in _grc_registry().items()}
    blocks = []
    for i, block in enumerate(graph.blocks):
        settings = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in getattr(block, "_grc_settings", {}).items()}
        blocks.append({"id": names.get(type(block), type(block).__name__) + "<complex64>", "parameters": {"name": getattr(block, "name", f"block{i}"), **settings}})
    label = {id(b): blocks[i]["parameters"]["name"] for i, b in enumerate(graph.blocks)}
    connections = [[label[id(s)], 0, label[id(d)], 0] + ([m] if m is not None else []) for s, d, m in graph.edges]
    return yaml.safe_dump({"blocks": blocks, "connections": connections}, sort_keys=False)
