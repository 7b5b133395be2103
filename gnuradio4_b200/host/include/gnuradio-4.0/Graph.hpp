// gr4b200 host layer -- gr::Graph: block list + edge list, same calls as the reference
// (core/include/gnuradio-4.0/Graph.hpp:425 emplaceBlock, :580/:596 connect, EdgeParameters BlockModel.hpp:64-72).
// Edge memory is decided when the scheduler connects pending edges: both ends on the same CUDA device => HBM ring
// (gr4b200_ring_*), both ends on the host => host ring, anything else is refused: domain transitions are explicit
// blocks (gr::cuda::H2D / D2H), the rule the reference states in core/README.md "Ports".
#pragma once

#include <algorithm>
#include <expected>
#include <memory>
#include <numeric>
#include <string>
#include <vector>

#include "Block.hpp"

namespace gr {

struct EdgeParameters {
    std::size_t minBufferSize = 65536; // items; the reference default for arithmetic types (Graph.hpp:102)
    std::int32_t weight       = 0;
    std::string  name         = "unnamed edge";
    std::string  domain       = "";    // where the edge's buffer lives when its two blocks do not agree (reference: EdgeParameters::domain,
                                       // Graph.hpp:709-762). "gpu:cuda:N" with N = the CONSUMER's device makes an edge between two GPUs a
                                       // ring in the consumer's HBM that the producer's kernels store into over NVLink (peer access).
};

struct Edge {
    BlockModel*    source;
    std::size_t    sourcePort;
    BlockModel*    destination;
    std::size_t    destinationPort;
    EdgeParameters parameters;
    std::shared_ptr<EdgeBuffer> buffer;
};

class Graph {
public:
    Graph()                        = default;
    Graph(Graph&&)                 = default;
    Graph& operator=(Graph&&)      = default;
    Graph(const Graph&)            = delete;
    Graph& operator=(const Graph&) = delete;

    template<typename TBlock>
    TBlock& emplaceBlock(property_map initialSettings = {}) {
        auto  wrapper = std::make_unique<BlockWrapper<TBlock>>(std::move(initialSettings));
        auto& ref     = wrapper->block();
        _blocks.push_back(std::move(wrapper));
        return ref;
    }

    // g.connect<"out", "in">(source, destination[, EdgeParameters{...}])
    template<meta::fixed_string SourcePort, meta::fixed_string DestinationPort, typename TSource, typename TDestination>
    std::expected<void, Error> connect(TSource& source, TDestination& destination, EdgeParameters parameters = {}) {
        return connect(source, std::string(SourcePort.view()), destination, std::string(DestinationPort.view()), std::move(parameters));
    }

    template<typename TSource, typename TDestination>
    std::expected<void, Error> connect(TSource& source, const std::string& sourcePort, TDestination& destination, const std::string& destinationPort, EdgeParameters parameters = {}) {
        BlockModel* src = find(&source);
        BlockModel* dst = find(&destination);
        if (src == nullptr || dst == nullptr) {
            return std::unexpected(Error{"connect: block is not part of this graph"});
        }
        const int sp = src->outputPortIndex(sourcePort), dp = dst->inputPortIndex(destinationPort);
        if (sp < 0) {
            return std::unexpected(Error{"connect: '" + std::string(src->typeName()) + "' has no output port '" + sourcePort + "'"});
        }
        if (dp < 0) {
            return std::unexpected(Error{"connect: '" + std::string(dst->typeName()) + "' has no input port '" + destinationPort + "'"});
        }
        if (src->outputItemBytes(static_cast<std::size_t>(sp)) != dst->inputItemBytes(static_cast<std::size_t>(dp))) {
            return std::unexpected(Error{"connect: port types differ in size"});
        }
        for (const auto& e : _edges) { // an output may feed several inputs (one writer, N readers); an input has one source
            if (e.destination == dst && e.destinationPort == static_cast<std::size_t>(dp)) {
                return std::unexpected(Error{"connect: input port already connected"});
            }
        }
        _edges.push_back(Edge{src, static_cast<std::size_t>(sp), dst, static_cast<std::size_t>(dp), std::move(parameters), nullptr});
        return {};
    }

    [[nodiscard]] std::vector<std::unique_ptr<BlockModel>>& blocks() noexcept { return _blocks; }
    [[nodiscard]] std::vector<Edge>&                        edges() noexcept { return _edges; }

    // reference: Graph::connectPendingEdges (Graph.hpp:840) -> applyEdgeConnection (:698-783)
    std::expected<void, Error> connectPendingEdges() {
        for (auto& e : _edges) {
            if (e.buffer) {
                continue;
            }
            const bool srcDevice = e.source->outputOnDevice(e.sourcePort), dstDevice = e.destination->inputOnDevice(e.destinationPort);
            if (srcDevice != dstDevice) {
                return std::unexpected(Error{"edge '" + std::string(e.source->name()) + "' -> '" + std::string(e.destination->name()) + "' crosses the host/device boundary: insert gr::cuda::H2D / gr::cuda::D2H"});
            }
            int device = 0;
            if (srcDevice) {
                const int a = e.source->outputDevice(e.sourcePort), b = e.destination->inputDevice(e.destinationPort);
                device = a;
                if (a != b) {
                    // the push model: the ring lives with the consumer, the producer's kernels write it through peer access
                    // (plain stores over NVLink; the ring's events order the two streams across the devices)
                    const auto domain = ComputeDomain::parse(e.parameters.domain);
                    if (e.parameters.domain.empty() || !domain.isCuda() || domain.deviceIndex != b) {
                        return std::unexpected(Error{"edge between different CUDA devices: insert gr::cuda::PeerCopy, or place the edge on the consumer's device (EdgeParameters{.domain = \"gpu:cuda:" + std::to_string(b) + "\"})"});
                    }
                    if (gr4b200_peer_enable(a, b) != GR4B200_OK) {
                        return std::unexpected(Error{std::string("edge between different CUDA devices: ") + gr4b200_last_error() + " -- insert gr::cuda::PeerCopy"});
                    }
                    device = b;
                }
            }
            // an output that already has a buffer (an earlier edge from the same port): this edge is one more reader of it
            Edge* sibling = nullptr;
            for (auto& other : _edges) {
                if (&other != &e && other.buffer && other.source == e.source && other.sourcePort == e.sourcePort) {
                    sibling = &other;
                    break;
                }
            }
            if (sibling != nullptr) {
                try {
                    e.buffer = sibling->buffer;
                    e.destination->bindInput(e.destinationPort, e.buffer, e.buffer->addReader());
                } catch (const std::exception& ex) {
                    return std::unexpected(Error{ex.what()});
                }
                continue;
            }
            // Spans never wrap (no double mapping in HBM): a consumer reads whole multiples of its input_chunk_size
            // starting at item 0, so a capacity that is a multiple of it (for every reader of this output, and of the
            // producer's output_chunk_size) always ends on a chunk boundary. The reference gets the same effect from its
            // mirrored mapping (CircularBuffer.hpp:382-409).
            std::size_t unit = std::max<std::size_t>(e.source->outputChunkSize(), 1), minItems = e.parameters.minBufferSize;
            for (const auto& other : _edges) {
                if (other.source == e.source && other.sourcePort == e.sourcePort) {
                    unit     = std::lcm(unit, std::max<std::size_t>(other.destination->inputChunkSize(), 1));
                    minItems = std::max(minItems, other.parameters.minBufferSize);
                }
            }
            const std::size_t capacity = (std::max(minItems, 2 * unit) + unit - 1) / unit * unit;
            // a host edge that a copy block reads or writes on a CUDA stream is pinned (asynchronous copies need it)
            const bool pinned = !srcDevice && (e.source->workDevice() >= 0 || e.destination->workDevice() >= 0);
            // a device edge with a single reader keeps the past items that reader asks for in front of every span
            std::size_t readers = 0, history = 0;
            for (const auto& other : _edges) {
                if (other.source == e.source && other.sourcePort == e.sourcePort) {
                    ++readers;
                    history = std::max(history, other.destination->inputHistoryItems(other.destinationPort));
                }
            }
            history = srcDevice && readers == 1 && history <= capacity ? history : 0;
            // the ring keeps the history behind the reader's cursor out of the writer's reach: half a buffer more, so that
            // the writer can still fill one half while the reader works on the other
            const std::size_t capacityItems = history > 0 ? (capacity + capacity / 2 + unit - 1) / unit * unit : capacity;
            try {
                e.buffer = std::make_shared<EdgeBuffer>(e.source->outputItemBytes(e.sourcePort), capacityItems, srcDevice, device, pinned, history);
            } catch (const std::exception& ex) {
                return std::unexpected(Error{ex.what()});
            }
            e.source->bindOutput(e.sourcePort, e.buffer);
            e.destination->bindInput(e.destinationPort, e.buffer);
        }
        return {};
    }

private:
    BlockModel* find(void* raw) {
        for (auto& b : _blocks) {
            if (b->raw() == raw) {
                return b.get();
            }
        }
        return nullptr;
    }
    std::vector<std::unique_ptr<BlockModel>> _blocks;
    std::vector<Edge>                        _edges;
};

} // namespace gr
