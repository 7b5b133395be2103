// gr4b200 host layer -- gr::scheduler::{Simple, BreadthFirst, DepthFirst}: exchange(Graph&&), runAndWait()
// (reference: core/include/gnuradio-4.0/Scheduler.hpp:394, :581; hot loop poolWorker :838-975 -> traverseBlockListOnce
// :718-736; the three schedulers differ only in the order of the block list, :1944-1951, :1983-2057, :2061-2125).
// One launcher thread; every device block's work chunk is an asynchronous launch on a CUDA stream of its device: one
// stream for kernels, one for host -> device copies, one for device -> host copies (gr::cuda::H2D / D2H), so that
// uploads, kernels and downloads of consecutive chunks overlap; ordering between blocks is stream order plus the edge
// events. The loop never waits for the GPU while any block can still issue work; when none can, it blocks on the oldest
// copy event an edge is waiting for (no spinning). The graph is done when every block reported DONE; any ERROR stops
// the run (Scheduler.hpp:726-735).
#pragma once

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <deque>
#include <expected>
#include <map>
#include <memory>
#include <set>
#include <utility>
#include <vector>

#include "Graph.hpp"

namespace gr::scheduler {

enum class ExecutionPolicy { singleThreaded, multiThreaded }; // multiThreaded: one launcher thread per CUDA device (+ one for host-only blocks)

template<ExecutionPolicy execution = ExecutionPolicy::singleThreaded>
class Simple {
public:
    Simple() = default;
    explicit Simple(Graph&& graph) { (void)exchange(std::move(graph)); }
    virtual ~Simple() {
        _graph.reset(); // rings before streams
        for (auto& [key, stream] : _streams) {
            gr4b200_stream_destroy(stream);
        }
    }

    std::size_t max_work_items = std::numeric_limits<std::size_t>::max(); // soft cap handed to every work() call (Scheduler.hpp:719-723)
    // kernel streams per device that blocks with independent work chunks rotate over (1: one stream, strict order).
    // Measured on the FIR -> FFT flowgraph (profiles/r02f_bm_flowgraph_streams.txt): rotation does not pay -- the launcher
    // thread spends more on the cross-stream events than the overlapping kernel tails give back -- so the default is 1.
    std::size_t compute_streams = 1;
    std::size_t host_threads    = 1; // multiThreaded: launcher threads the host-only blocks are dealt out to (device blocks: one thread per device)

    std::expected<Graph*, Error> exchange(Graph&& graph) {
        _graph = std::make_unique<Graph>(std::move(graph));
        return _graph.get();
    }

    [[nodiscard]] Graph& graph() { return *_graph; }

    // the blocks in the order one pass of the work loop visits them (valid after runAndWait started; tests read it)
    [[nodiscard]] const std::vector<BlockModel*>& executionOrder() const noexcept { return _order; }

    // INITIALISED + the start of RUNNING in the reference's life cycle (Scheduler.hpp:746-809): blocks initialised, edges
    // allocated, streams created, start() hooks run. runAndWait() does it on demand; calling it beforehand keeps the
    // allocations out of a timed run.
    std::expected<void, Error> init() {
        if (_initialised) {
            return {};
        }
        if (!_graph) {
            return std::unexpected(Error{"no graph"});
        }
        try {
            for (auto& block : _graph->blocks()) {
                block->init();
            }
        } catch (const std::exception& ex) {
            return std::unexpected(Error{ex.what()});
        }
        if (auto connected = _graph->connectPendingEdges(); !connected) {
            return connected;
        }
        // every block that touches a device (device blocks, and the bridges H2D / D2H / PeerCopy) gets a stream of the
        // device its work is issued on -- per device one for kernels and one per copy direction; stream order replaces the
        // reference's thread order
        std::set<int> devices;
        try {
            for (auto& block : _graph->blocks()) {
                const int device = block->workDevice();
                if (device >= 0) {
                    block->setStream(streamFor(device, block->streamRole()));
                    devices.insert(device);
                    if (block->streamRole() == 0 && compute_streams > 1 && block->chunksIndependent()) {
                        // streams of their own (roles >= 16): the device's base kernel stream also carries the blocks that keep
                        // strict order (sources, sinks, the mixer) and the waits they issue; a rotating chunk queued behind
                        // those would wait for work it does not depend on
                        std::vector<void*> rotation;
                        for (std::size_t k = 0; k < compute_streams; ++k) {
                            rotation.push_back(streamFor(device, 16 + static_cast<int>(k)));
                        }
                        block->setStreams(std::move(rotation));
                    }
                }
            }
        } catch (const std::exception& ex) {
            return std::unexpected(Error{ex.what()});
        }
        _order = orderBlocks(*_graph);
        try {
            for (BlockModel* block : _order) {
                if (const int device = block->workDevice(); device >= 0) {
                    gr4b200_init(device);
                }
                block->start();
            }
        } catch (const std::exception& ex) {
            return std::unexpected(Error{ex.what()});
        }
        _severalDevices = devices.size() > 1;
        _initialised    = true;
        return {};
    }

    std::expected<void, Error> runAndWait() {
        if (auto ready = init(); !ready) {
            return ready;
        }
        if constexpr (execution == ExecutionPolicy::multiThreaded) {
            return runOnThreads();
        }
        int currentDevice = -1;
        std::size_t idleRounds = 0;
        // GR4B200_SCHED_PROFILE=1: host time the launcher thread spends inside every block's work() (the reference's
        // profiler hooks, Profiler.hpp, reduced to the one figure that decides small-chunk throughput here)
        static const bool profile = [] { const char* e = std::getenv("GR4B200_SCHED_PROFILE"); return e != nullptr && e[0] == '1'; }();
        std::vector<std::pair<double, std::size_t>> spent(_order.size(), {0.0, 0});
        const auto                                  runStart = std::chrono::steady_clock::now();
        while (true) {
            std::size_t done = 0, progressed = 0, sinks = 0, sinksDone = 0;
            for (std::size_t blockIndex = 0; blockIndex < _order.size(); ++blockIndex) {
                BlockModel* block = _order[blockIndex];
                if (const int device = block->workDevice(); device >= 0 && device != currentDevice && _severalDevices) {
                    gr4b200_init(device); // several devices in one graph: launches go to the calling thread's current device
                    currentDevice = device;
                }
                const auto         workStart = profile ? std::chrono::steady_clock::now() : std::chrono::steady_clock::time_point{};
                const work::Result result    = block->work(max_work_items);
                if (profile) {
                    spent[blockIndex].first += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - workStart).count();
                    spent[blockIndex].second += result.performed_work > 0 ? 1 : 0;
                }
                if (result.status == work::Status::ERROR) {
                    return std::unexpected(Error{"block '" + std::string(block->name()) + "' reported ERROR: " + gr4b200_last_error()});
                }
                done += result.status == work::Status::DONE ? 1 : 0;
                progressed += result.performed_work > 0 ? 1 : 0;
                if (block->outputCount() == 0) {
                    ++sinks;
                    sinksDone += result.status == work::Status::DONE ? 1 : 0;
                }
            }
            // every block finished, or every sink asked to stop (CountingSink's n_samples_max, NullSources.hpp:200-247)
            if (done == _order.size() || (sinks > 0 && sinksDone == sinks)) {
                break;
            }
            if (progressed > 0) {
                idleRounds = 0;
                continue;
            }
            // nobody could move: if an edge is waiting for a copy, block on that copy's event; if nothing is in flight the
            // graph cannot make progress any more (the reference's watchdog would warn here, Scheduler.hpp:977-1009)
            void* pending = nullptr;
            for (auto& edge : _graph->edges()) {
                if (edge.buffer && (pending = edge.buffer->oldestPendingEvent()) != nullptr) {
                    break;
                }
            }
            if (pending != nullptr) {
                if (gr4b200_event_synchronize(pending) != GR4B200_OK) {
                    return std::unexpected(Error{gr4b200_last_error()});
                }
                idleRounds = 0;
            } else if (++idleRounds > 1000) {
                return std::unexpected(Error{"flowgraph stalled: no block made progress"});
            }
        }
        const auto issueEnd = std::chrono::steady_clock::now();
        for (auto& [key, stream] : _streams) {
            if (gr4b200_stream_synchronize(stream) != GR4B200_OK) {
                return std::unexpected(Error{gr4b200_last_error()});
            }
        }
        if (profile) {
            std::fprintf(stderr, "[sched profile] issue loop %.1f us, drain %.1f us\n", std::chrono::duration<double, std::micro>(issueEnd - runStart).count(), std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - issueEnd).count());
            for (std::size_t k = 0; k < _order.size(); ++k) {
                std::fprintf(stderr, "[sched profile]   %-32s %10.1f us in work(), %8zu productive calls, %.2f us each\n", std::string(_order[k]->name()).c_str(), spent[k].first, spent[k].second, spent[k].second > 0 ? spent[k].first / static_cast<double>(spent[k].second) : 0.0);
            }
        }
        return {};
    }

private:
    // ExecutionPolicy::multiThreaded (reference: Scheduler.hpp:1929-1968 deals the block list out to pool threads). Here the
    // natural partition is the device: one launcher thread per CUDA device (it binds the device once and issues every
    // launch of that device's blocks), one more for the blocks that never touch a device. Edges whose two ends live on
    // different threads are the hand-off queues, exactly as the reference's CircularBuffers are between its job lists;
    // each thread runs the single-threaded loop over its own list.
    std::expected<void, Error> runOnThreads() {
        std::map<int, std::vector<BlockModel*>> lists;
        std::size_t                              totalSinks = 0;
        std::size_t hostBlocks = 0;
        for (BlockModel* block : _order) {
            // device blocks: the list of their device; host-only blocks: dealt round-robin over `host_threads` lists (keys < 0)
            const int device = block->workDevice();
            const int key    = device >= 0 ? device : -1 - static_cast<int>(hostBlocks++ % std::max<std::size_t>(host_threads, 1));
            lists[key].push_back(block);
            totalSinks += block->outputCount() == 0 ? 1 : 0;
        }
        std::atomic<bool>        stop{false};
        std::atomic<std::size_t> sinksDone{0}, progress{0};
        std::mutex               errorMutex;
        std::string              errorMessage;
        auto                     worker = [&](int device, std::vector<BlockModel*> list) {
            if (device >= 0) {
                gr4b200_init(device);
            }
            std::vector<bool> finished(list.size(), false);
            auto              lastProgress = std::chrono::steady_clock::now();
            std::size_t       seenProgress = progress.load();
            while (!stop.load()) {
                std::size_t done = 0, progressed = 0;
                for (std::size_t k = 0; k < list.size(); ++k) {
                    const work::Result result = list[k]->work(max_work_items);
                    if (result.status == work::Status::ERROR) {
                        const std::lock_guard<std::mutex> lock(errorMutex);
                        errorMessage = "block '" + std::string(list[k]->name()) + "' reported ERROR: " + gr4b200_last_error();
                        stop         = true;
                        return;
                    }
                    if (result.status == work::Status::DONE) {
                        ++done;
                        if (!finished[k]) {
                            finished[k] = true;
                            if (list[k]->outputCount() == 0 && sinksDone.fetch_add(1) + 1 == totalSinks) {
                                stop = true; // every sink asked to stop (CountingSink's n_samples_max)
                            }
                        }
                    }
                    progressed += result.performed_work > 0 ? 1 : 0;
                }
                if (done == list.size()) {
                    return;
                }
                if (progressed > 0) {
                    progress.fetch_add(1);
                    continue;
                }
                // nothing to do right now: a copy this thread's edges wait for, or work of another thread
                void* pending = nullptr;
                for (auto& edge : _graph->edges()) {
                    const bool mine = std::find(list.begin(), list.end(), edge.source) != list.end() || std::find(list.begin(), list.end(), edge.destination) != list.end();
                    if (mine && edge.buffer && (pending = edge.buffer->oldestPendingEvent()) != nullptr) {
                        break;
                    }
                }
                if (pending != nullptr) {
                    gr4b200_event_synchronize(pending);
                    continue;
                }
                std::this_thread::yield();
                const auto now = std::chrono::steady_clock::now();
                if (const std::size_t p = progress.load(); p != seenProgress) {
                    seenProgress = p;
                    lastProgress = now;
                } else if (now - lastProgress > std::chrono::seconds(10)) { // the reference's watchdog would warn here (Scheduler.hpp:977-1009)
                    const std::lock_guard<std::mutex> lock(errorMutex);
                    if (errorMessage.empty()) {
                        errorMessage = "flowgraph stalled: no block made progress for 10 s";
                    }
                    stop = true;
                    return;
                }
            }
        };
        std::vector<std::thread> threads;
        for (auto& [device, list] : lists) {
            threads.emplace_back(worker, device, list);
        }
        for (auto& thread : threads) {
            thread.join();
        }
        if (!errorMessage.empty()) {
            return std::unexpected(Error{errorMessage});
        }
        for (auto& [key, stream] : _streams) {
            if (key.first >= 0) {
                gr4b200_init(key.first);
            }
            if (gr4b200_stream_synchronize(stream) != GR4B200_OK) {
                return std::unexpected(Error{gr4b200_last_error()});
            }
        }
        return {};
    }

protected:
    // Simple: the order the blocks were emplaced in (Scheduler.hpp:1944-1951)
    virtual std::vector<BlockModel*> orderBlocks(Graph& graph) const {
        std::vector<BlockModel*> order;
        for (auto& block : graph.blocks()) {
            order.push_back(block.get());
        }
        return order;
    }

    // blocks without an incoming edge, in emplace order, and every block's successors sorted by output port, then by the
    // order the edges were connected in (graph::computeAdjacencyList / findSourceBlocks)
    struct Topology {
        std::vector<BlockModel*>                         sources;
        std::map<BlockModel*, std::vector<BlockModel*>> successors;
    };
    static Topology topologyOf(Graph& graph) {
        Topology              t;
        std::set<BlockModel*> hasInput;
        std::vector<Edge*>    edges;
        for (auto& edge : graph.edges()) {
            hasInput.insert(edge.destination);
            edges.push_back(&edge);
        }
        std::stable_sort(edges.begin(), edges.end(), [](const Edge* a, const Edge* b) { return a->sourcePort < b->sourcePort; });
        for (const Edge* edge : edges) {
            t.successors[edge->source].push_back(edge->destination);
        }
        for (auto& block : graph.blocks()) {
            if (!hasInput.contains(block.get())) {
                t.sources.push_back(block.get());
            }
        }
        return t;
    }

private:
    void* streamFor(int device, int role) {
        const auto key = std::make_pair(device, role);
        auto       it  = _streams.find(key);
        if (it != _streams.end()) {
            return it->second;
        }
        if (gr4b200_init(device) != GR4B200_OK) {
            throw exception(std::string("cannot initialise CUDA device: ") + gr4b200_last_error());
        }
        void* stream = gr4b200_stream_create();
        if (stream == nullptr) {
            throw exception(std::string("cannot create CUDA stream: ") + gr4b200_last_error());
        }
        _streams.emplace(key, stream);
        return stream;
    }

    std::unique_ptr<Graph>   _graph;
    std::map<std::pair<int, int>, void*> _streams; // (device, role) -> stream
    std::vector<BlockModel*> _order;
    bool                     _initialised    = false;
    bool                     _severalDevices = false;
};

// breadth-first from the source blocks: a block is queued the first time an edge reaches it (Scheduler.hpp:1983-2057)
template<ExecutionPolicy execution = ExecutionPolicy::singleThreaded>
class BreadthFirst : public Simple<execution> {
public:
    using Simple<execution>::Simple;

protected:
    std::vector<BlockModel*> orderBlocks(Graph& graph) const override {
        const auto               topology = Simple<execution>::topologyOf(graph);
        std::vector<BlockModel*> order;
        std::set<BlockModel*>    reached(topology.sources.begin(), topology.sources.end());
        std::deque<BlockModel*>  queue(topology.sources.begin(), topology.sources.end());
        while (!queue.empty()) {
            BlockModel* current = queue.front();
            queue.pop_front();
            order.push_back(current);
            if (auto it = topology.successors.find(current); it != topology.successors.end()) {
                for (BlockModel* next : it->second) {
                    if (reached.insert(next).second) {
                        queue.push_back(next);
                    }
                }
            }
        }
        return order;
    }
};

// depth-first from every source block in turn (Scheduler.hpp:2061-2125)
template<ExecutionPolicy execution = ExecutionPolicy::singleThreaded>
class DepthFirst : public Simple<execution> {
public:
    using Simple<execution>::Simple;

protected:
    std::vector<BlockModel*> orderBlocks(Graph& graph) const override {
        const auto               topology = Simple<execution>::topologyOf(graph);
        std::vector<BlockModel*> order;
        std::set<BlockModel*>    visited;
        std::vector<BlockModel*> stack;
        for (BlockModel* source : topology.sources) {
            stack.push_back(source);
            while (!stack.empty()) {
                BlockModel* current = stack.back();
                stack.pop_back();
                if (!visited.insert(current).second) {
                    continue;
                }
                order.push_back(current);
                if (auto it = topology.successors.find(current); it != topology.successors.end()) {
                    for (auto next = it->second.rbegin(); next != it->second.rend(); ++next) { // first successor on top
                        stack.push_back(*next);
                    }
                }
            }
        }
        return order;
    }
};

} // namespace gr::scheduler
