// gr4b200 host layer -- gr::Block<Derived>, ports, work(): the block-author surface of GNU Radio 4 on top of the
// B200 engine. A block written against the reference compiles unchanged against this header as long as it stays inside
// the surface listed in SURVEY.md 8(b):
//   struct X : gr::Block<X [, gr::Resampling<I, O, isConst>]> { PortIn<T> in; PortOut<T> out; <settings>;
//       GR_MAKE_REFLECTABLE(X, in, out, ...);  one of processOne / processBulk;  optional settingsChanged(old, new) };
// New, opt-in, for device execution (the analogue of the reference's inert processBulk_sycl hook,
// core/include/gnuradio-4.0/BlockTraits.hpp:422):
//   work::Status processBulk_cuda(void* stream, const TIn* in, TOut* out, std::size_t nIn, std::size_t nOut);
// A block runs on the device iff it has that member AND its `compute_domain` setting parses to gpu:cuda[:N]
// (reference seam: Block.hpp:1855-1862). Device blocks without a host body refuse a host compute_domain at init -- there is
// no silent CPU fallback for the accelerated path.
//
// What work() reproduces from the reference (Block.hpp:2028-2172): samples to process = min over ports of
// available / free space, floored to whole input_chunk_size multiples with output = k * output_chunk_size
// (:1610-1635), capped by the port's max_samples and the scheduler's requested work; the whole chunk is consumed and
// published; DONE propagates downstream once an input edge is drained and its producer has finished.
#pragma once

#include <cstddef>
#include <cstdint>
#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>
#include <deque>
#include <limits>
#include <memory>
#include <span>
#include <stdexcept>
#include <string>
#include <string_view>
#include <tuple>
#include <type_traits>
#include <vector>

#include "../../../../include/gr4b200.h"
#include "ComputeDomain.hpp"
#include "Value.hpp"
#include "meta.hpp"

namespace gr {

namespace work {
enum class Status { ERROR = -100, INSUFFICIENT_OUTPUT_ITEMS = -3, INSUFFICIENT_INPUT_ITEMS = -2, DONE = -1, OK = 0 }; // WorkStatus.hpp:12-18
struct Result {
    std::size_t requested_work = std::numeric_limits<std::size_t>::max();
    std::size_t performed_work = 0;
    Status      status         = Status::OK;
};
} // namespace work

struct exception : std::runtime_error {
    using std::runtime_error::runtime_error;
};

struct Error {
    std::string message;
};

// ---- edges ---------------------------------------------------------------------------------------------------------
// ---- tags ------------------------------------------------------------------------------------------------------------
// Stream tags (core/include/gnuradio-4.0/Tag.hpp): a property_map attached to one sample of a stream. Here `index` is
// the ABSOLUTE sample index on the edge that carries the tag. Tags are host-side metadata even when the samples live in
// HBM; the scheduling rules are the reference's: a tag always travels with the FIRST sample of a work chunk (a later tag
// in the available range ends the chunk in front of it, Block.hpp:1960-1971), keys that name a setting of the consuming
// block update that setting (settings auto-update), and tags are forwarded to every output with `sample_rate`
// rescaled by output_chunk_size / input_chunk_size on resampling blocks (Block.hpp:1088-1100).
struct Tag {
    std::size_t  index = 0;
    property_map map;
};
namespace tag {
inline constexpr std::string_view kPrefix     = "gr:"; // wire prefix of the default tags (Tag.hpp GR_TAG_PREFIX)
inline constexpr std::string_view SAMPLE_RATE = "sample_rate";
// bare setting name of a tag key ("gr:sample_rate" -> "sample_rate")
inline std::string_view settingsKey(std::string_view wireKey) { return wireKey.starts_with(kPrefix) ? wireKey.substr(kPrefix.size()) : wireKey; }
// the settings every block forwards downstream as a tag when they change (Tag.hpp kDefaultTags, stream-related subset)
inline bool isDefaultTag(std::string_view bareKey) { return bareKey == "sample_rate" || bareKey == "signal_name" || bareKey == "signal_unit" || bareKey == "signal_min" || bareKey == "signal_max"; }
} // namespace tag

// One producer, N consumers, contiguous spans only (see include/gr4b200.h "HBM edge ring"). The host flavour keeps the
// same cursor protocol over host memory so that host-only graphs (BASELINE config #1) run without a GPU.
// A host edge next to a block that works on a CUDA stream (gr::cuda::H2D reads it, gr::cuda::D2H writes it) is PINNED
// and its cursors follow the stream: `consume(n, stream)` / `publish(n, stream)` record an event behind the copy that was
// just enqueued and the cursor only becomes visible to the other side once that event has completed (polled, never
// waited for, in available() / writable()). The copy blocks therefore never synchronise: the launcher thread keeps
// issuing work while copies are in flight, and the host consumer sees a span exactly when its bytes have landed.
class EdgeBuffer {
public:
    // historyItems (device edges with one reader): that many items in front of every span the reader gets stay valid --
    // the stream's own past, zeros before its start (gr4b200_ring_create's history)
    EdgeBuffer(std::size_t itemBytes, std::size_t capacityItems, bool onDevice, int device, bool pinnedHost = false, std::size_t historyItems = 0) : _itemBytes(itemBytes), _capacity(capacityItems * itemBytes), _onDevice(onDevice), _pinned(pinnedHost && !onDevice), _historyItems(onDevice ? historyItems : 0) {
        if (onDevice) {
            _ring = gr4b200_ring_create(device, _capacity, _historyItems * itemBytes);
            if (_ring == nullptr) {
                throw exception(std::string("device edge: ") + gr4b200_last_error());
            }
        } else if (_pinned) {
            _hostBase = static_cast<std::byte*>(gr4b200_malloc_host(_capacity));
            if (_hostBase == nullptr) {
                throw exception(std::string("pinned host edge: ") + gr4b200_last_error());
            }
        } else {
            _host.resize(_capacity);
            _hostBase = _host.data();
        }
    }
    ~EdgeBuffer() {
        if (_ring != nullptr) {
            gr4b200_ring_destroy(_ring);
        }
        _pendingPublish.destroy();
        for (auto& queue : _pendingConsume) {
            queue.destroy();
        }
        if (_pinned && _hostBase != nullptr) {
            gr4b200_free_host(_hostBase);
        }
        for (auto& scratch : _linear) {
            if (scratch.device != nullptr) {
                gr4b200_free(scratch.device);
            }
        }
    }
    EdgeBuffer(const EdgeBuffer&)            = delete;
    EdgeBuffer& operator=(const EdgeBuffer&) = delete;

    [[nodiscard]] bool        onDevice() const noexcept { return _onDevice; }
    [[nodiscard]] bool        pinned() const noexcept { return _pinned; }
    [[nodiscard]] std::size_t historyItems() const noexcept { return _historyItems; }
    [[nodiscard]] std::size_t itemBytes() const noexcept { return _itemBytes; }
    // one writer, N readers (CircularBuffer.hpp:476-477): reader 0 exists from the start, more join before data flows
    int addReader() {
        const std::lock_guard<std::recursive_mutex> lock(_mutex);
        if (_onDevice) {
            const int reader = gr4b200_ring_add_reader(_ring);
            if (reader < 0) {
                throw exception(std::string("device edge: ") + gr4b200_last_error());
            }
            _itemsConsumed.push_back(0);
            return reader;
        }
        _consumed.push_back(0);
        _consumedIssued.push_back(0);
        _pendingConsume.emplace_back();
        _itemsConsumed.push_back(0);
        return static_cast<int>(_consumed.size()) - 1;
    }
    [[nodiscard]] std::size_t available(int reader = 0) { // items published and contiguous for this reader
        const std::lock_guard<std::recursive_mutex> lock(_mutex);
        if (_onDevice) {
            return gr4b200_ring_available_for(_ring, reader) / _itemBytes;
        }
        poll();
        const std::uint64_t consumed = _consumedIssued[static_cast<std::size_t>(reader)];
        const std::size_t   pending = static_cast<std::size_t>(_written - consumed), contiguous = _capacity - static_cast<std::size_t>(consumed % _capacity);
        return std::min(pending, contiguous) / _itemBytes;
    }
    [[nodiscard]] std::size_t writable() {
        const std::lock_guard<std::recursive_mutex> lock(_mutex);
        if (_onDevice) {
            return gr4b200_ring_writable(_ring) / _itemBytes;
        }
        poll();
        const std::uint64_t slowest   = *std::min_element(_consumed.begin(), _consumed.end());
        const std::size_t   freeBytes = _capacity - static_cast<std::size_t>(_writtenIssued - slowest), contiguous = _capacity - static_cast<std::size_t>(_writtenIssued % _capacity);
        return std::min(freeBytes, contiguous) / _itemBytes;
    }
    void* reserve(std::size_t items, void* stream) {
        const std::lock_guard<std::recursive_mutex> lock(_mutex);
        if (_onDevice) {
            return gr4b200_ring_reserve(_ring, items * _itemBytes, stream);
        }
        return items <= writable() ? _hostBase + _writtenIssued % _capacity : nullptr;
    }
    void publish(std::size_t items, void* stream) {
        const std::lock_guard<std::recursive_mutex> lock(_mutex);
        if (_onDevice) {
            if (gr4b200_ring_publish(_ring, items * _itemBytes, stream) != GR4B200_OK) {
                failed = true;
            }
        } else {
            _writtenIssued += items * _itemBytes;
            if (stream != nullptr && items > 0) { // written by a copy that is still in flight on `stream`
                defer(_pendingPublish, _writtenIssued, stream);
            } else if (_pendingPublish.pending.empty()) {
                _written = _writtenIssued;
            } else {
                _pendingPublish.pending.back().cursor = _writtenIssued; // rides on the last copy still in flight
            }
        }
        _itemsPublished += items;
    }
    // tag on the sample `offset` items behind everything published so far (i.e. inside the chunk about to be published)
    void publishTag(property_map map, std::size_t offset = 0) {
        const std::lock_guard<std::recursive_mutex> lock(_mutex);
        if (map.empty()) {
            return;
        }
        const std::size_t index = static_cast<std::size_t>(_itemsPublished) + offset;
        if (!tags.empty() && tags.back().index == index) { // same sample: merge, later keys win
            for (auto& [key, value] : map) {
                tags.back().map.insert_or_assign(key, std::move(value));
            }
        } else {
            tags.push_back(Tag{index, std::move(map)});
        }
    }
    [[nodiscard]] std::size_t itemsPublished() {
        const std::lock_guard<std::recursive_mutex> lock(_mutex);
        return static_cast<std::size_t>(_itemsPublished);
    }
    [[nodiscard]] std::size_t itemsConsumed(int reader = 0) {
        const std::lock_guard<std::recursive_mutex> lock(_mutex);
        return static_cast<std::size_t>(_itemsConsumed[static_cast<std::size_t>(reader)]);
    }
    // the writer and every reader of an edge may live on different launcher threads (scheduler::ExecutionPolicy::
    // multiThreaded): each cursor has one owner, the mutex makes the other side's view of it (and the tag queue) coherent
    [[nodiscard]] std::recursive_mutex& mutex() noexcept { return _mutex; }
    std::deque<Tag> tags; // ascending index; dropped once every reader has passed them
    const void* get(std::size_t items, void* stream, int reader = 0) {
        const std::lock_guard<std::recursive_mutex> lock(_mutex);
        if (_onDevice) {
            return gr4b200_ring_get_for(_ring, reader, items * _itemBytes, stream);
        }
        return items <= available(reader) ? _hostBase + _consumedIssued[static_cast<std::size_t>(reader)] % _capacity : nullptr;
    }
    void consume(std::size_t items, void* stream, int reader = 0) {
        const std::lock_guard<std::recursive_mutex> lock(_mutex);
        const auto r = static_cast<std::size_t>(reader);
        if (_onDevice) {
            if (gr4b200_ring_consume_for(_ring, reader, items * _itemBytes, stream) != GR4B200_OK) {
                failed = true;
            }
        } else {
            _consumedIssued[r] += items * _itemBytes;
            if (stream != nullptr && items > 0) { // read by a copy that is still in flight on `stream`
                defer(_pendingConsume[r], _consumedIssued[r], stream);
            } else if (_pendingConsume[r].pending.empty()) {
                _consumed[r] = _consumedIssued[r];
            } else {
                _pendingConsume[r].pending.back().cursor = _consumedIssued[r];
            }
        }
        _itemsConsumed[r] += items;
        const std::uint64_t slowest = *std::min_element(_itemsConsumed.begin(), _itemsConsumed.end());
        while (!tags.empty() && tags.front().index < slowest) {
            tags.pop_front();
        }
    }
    // items published for this reader, whether or not they are contiguous in the ring
    [[nodiscard]] std::size_t pending(int reader = 0) {
        const std::lock_guard<std::recursive_mutex> lock(_mutex);
        if (_onDevice) {
            return gr4b200_ring_pending_for(_ring, reader) / _itemBytes;
        }
        poll();
        return static_cast<std::size_t>(_written - _consumedIssued[static_cast<std::size_t>(reader)]) / _itemBytes;
    }
    // `items` published items as ONE span even where they run across the end of the ring: readers that do not consume whole
    // chunks (Stride<> with overlap) meet that case; the reference's ring is mapped twice for it (CircularBuffer.hpp:75-173),
    // here the two pieces are copied into a scratch span of the reader (host: memcpy, device: two copies on `stream`)
    const void* getLinear(std::size_t items, void* stream, int reader = 0) {
        const std::lock_guard<std::recursive_mutex> lock(_mutex);
        if (items <= available(reader)) {
            return get(items, stream, reader);
        }
        if (items > pending(reader)) {
            return nullptr;
        }
        const std::size_t bytes = items * _itemBytes;
        if (_linear.size() <= static_cast<std::size_t>(reader)) {
            _linear.resize(static_cast<std::size_t>(reader) + 1);
        }
        Linear& scratch = _linear[static_cast<std::size_t>(reader)];
        if (_onDevice) {
            if (scratch.deviceBytes < bytes) {
                if (scratch.device != nullptr) {
                    gr4b200_stream_synchronize(stream);
                    gr4b200_free(scratch.device);
                }
                scratch.device      = gr4b200_malloc(bytes);
                scratch.deviceBytes = scratch.device != nullptr ? bytes : 0;
            }
            if (scratch.device == nullptr || gr4b200_ring_read_for(_ring, reader, bytes, scratch.device, stream) != GR4B200_OK) {
                failed = true;
                return nullptr;
            }
            return scratch.device;
        }
        scratch.host.resize(bytes);
        const std::size_t begin = static_cast<std::size_t>(_consumedIssued[static_cast<std::size_t>(reader)] % _capacity);
        const std::size_t first = std::min(bytes, _capacity - begin);
        std::memcpy(scratch.host.data(), _hostBase + begin, first);
        std::memcpy(scratch.host.data() + first, _hostBase, bytes - first);
        return scratch.host.data();
    }
    // spans whose bytes are still travelling: the consumer must not take the edge for drained yet
    [[nodiscard]] bool publishPending() {
        const std::lock_guard<std::recursive_mutex> lock(_mutex);
        poll();
        return !_pendingPublish.pending.empty();
    }
    // the oldest event any cursor of this edge is waiting for, or nullptr: the scheduler blocks on it when no block can
    // make progress instead of spinning
    [[nodiscard]] void* oldestPendingEvent() {
        const std::lock_guard<std::recursive_mutex> lock(_mutex);
        poll();
        if (!_pendingPublish.pending.empty()) {
            return _pendingPublish.pending.front().event;
        }
        for (auto& queue : _pendingConsume) {
            if (!queue.pending.empty()) {
                return queue.pending.front().event;
            }
        }
        return nullptr;
    }
    std::atomic<bool> producerDone{false};
    std::atomic<bool> failed{false}; // a cursor operation on the device ring reported an error (gr4b200_last_error has the reason)

private:
    struct Pending {
        std::uint64_t cursor; // the cursor value that becomes visible when `event` has completed
        void*         event;
    };
    // one cursor's in-flight updates; the events are recorded by one block, i.e. always on the same device, and recycled
    struct CursorQueue {
        std::deque<Pending> pending;
        std::vector<void*>  pool;
        void destroy() {
            for (auto& p : pending) {
                gr4b200_event_destroy(p.event);
            }
            for (void* e : pool) {
                gr4b200_event_destroy(e);
            }
            pending.clear();
            pool.clear();
        }
    };
    void defer(CursorQueue& queue, std::uint64_t cursor, void* stream) {
        void* event = nullptr;
        if (!queue.pool.empty()) {
            event = queue.pool.back();
            queue.pool.pop_back();
        } else {
            event = gr4b200_event_create();
        }
        if (event == nullptr || gr4b200_event_record(event, stream) != GR4B200_OK) {
            failed = true;
            if (event != nullptr) {
                queue.pool.push_back(event);
            }
            return;
        }
        queue.pending.push_back(Pending{cursor, event});
    }
    void drain(CursorQueue& queue, std::uint64_t& visible) {
        while (!queue.pending.empty()) {
            const int state = gr4b200_event_query(queue.pending.front().event);
            if (state == 0) {
                break;
            }
            if (state < 0) {
                failed = true;
            }
            visible = queue.pending.front().cursor;
            queue.pool.push_back(queue.pending.front().event);
            queue.pending.pop_front();
        }
    }
    void poll() {
        drain(_pendingPublish, _written);
        for (std::size_t r = 0; r < _pendingConsume.size(); ++r) {
            drain(_pendingConsume[r], _consumed[r]);
        }
    }

    std::recursive_mutex   _mutex;
    std::size_t            _itemBytes;
    std::size_t            _capacity;
    bool                   _onDevice;
    bool                   _pinned;
    std::size_t            _historyItems = 0;
    gr4b200_ring*          _ring = nullptr;
    std::vector<std::byte> _host;
    std::byte*             _hostBase = nullptr;
    // host ring cursors in bytes: `issued` moves when a block calls publish / consume, the visible one when the copy
    // behind it (if any) has completed
    std::uint64_t              _written = 0, _writtenIssued = 0;
    std::vector<std::uint64_t> _consumed{0}, _consumedIssued{0};
    CursorQueue                _pendingPublish;
    std::vector<CursorQueue>   _pendingConsume{1};
    struct Linear { // per reader: where a span that wraps is put together
        std::vector<std::byte> host;
        void*                  device      = nullptr;
        std::size_t            deviceBytes = 0;
    };
    std::vector<Linear>        _linear;
    std::uint64_t              _itemsPublished = 0;     // items since stream start (tag positions)
    std::vector<std::uint64_t> _itemsConsumed{0};       // per reader
};

// ---- ports -----------------------------------------------------------------------------------------------------------
enum class PortDirection { INPUT, OUTPUT };

template<typename T, PortDirection Dir>
struct Port {
    using value_type                         = T;
    static constexpr PortDirection direction = Dir;
    std::size_t                    min_samples = 1;
    std::size_t                    max_samples = std::numeric_limits<std::size_t>::max();
    std::shared_ptr<EdgeBuffer>    edge;       // set by Graph::connect / the scheduler
    int                            reader = 0; // input ports: which of the edge's readers this port is
};
template<typename T>
using PortIn = Port<T, PortDirection::INPUT>;
template<typename T>
using PortOut = Port<T, PortDirection::OUTPUT>;

template<typename T>
struct is_port : std::false_type {};
template<typename T, PortDirection D>
struct is_port<Port<T, D>> : std::true_type {};
// a dynamic port collection, `std::vector<gr::PortIn<T>> in;` (reference: Port.hpp port collections, named "in#0", "in#1", ...)
template<typename T>
struct is_port_vector : std::false_type {};
template<typename T, PortDirection D>
struct is_port_vector<std::vector<Port<T, D>>> : std::true_type {};

// ---- block-level attributes ------------------------------------------------------------------------------------------
template<std::size_t InputChunk = 1, std::size_t OutputChunk = 1, bool IsConst = false>
struct Resampling { // annotated.hpp:121-162
    static constexpr std::size_t kInputChunkSize  = InputChunk;
    static constexpr std::size_t kOutputChunkSize = OutputChunk;
    static constexpr bool        kIsConst         = IsConst;
};

template<std::uint64_t StrideValue = 0, bool IsConst = false>
struct Stride { // annotated.hpp:150-162: samples between the starts of consecutive chunks; < N overlaps, > N skips, 0 = back to back
    static constexpr std::size_t kStride  = StrideValue;
    static constexpr bool        kIsConst = IsConst;
    static constexpr bool        kEnabled = !IsConst || StrideValue > 0;
};

// `Annotated<float, "sample rate", ...> sample_rate = 1.f;` -- the description arguments are accepted and ignored
template<typename T, meta::fixed_string Description = "", typename... Attributes>
struct Annotated {
    using value_type = T;
    T value{};
    constexpr Annotated() = default;
    constexpr Annotated(const T& v) : value(v) {}
    constexpr Annotated& operator=(const T& v) {
        value = v;
        return *this;
    }
    constexpr operator T&() noexcept { return value; }
    constexpr operator const T&() const noexcept { return value; }
};
template<meta::fixed_string>
struct Doc {};
template<meta::fixed_string>
struct Unit {};
struct Visible {};
template<auto Lo, auto Hi>
struct Limits {};

template<typename T>
struct is_annotated : std::false_type {};
template<typename T, meta::fixed_string D, typename... A>
struct is_annotated<Annotated<T, D, A...>> : std::true_type {};

namespace detail {
template<typename T>
struct ResamplingOf {
    using type = Resampling<1, 1, true>;
};
template<std::size_t I, std::size_t O, bool C>
struct ResamplingOf<Resampling<I, O, C>> {
    using type = Resampling<I, O, C>;
};
template<typename... Args>
struct FirstResampling {
    using type = Resampling<1, 1, true>;
};
template<typename A, typename... Rest>
struct FirstResampling<A, Rest...> {
    using type = std::conditional_t<requires { A::kInputChunkSize; }, typename ResamplingOf<A>::type, typename FirstResampling<Rest...>::type>;
};
template<typename... Args>
struct FirstStride {
    using type = Stride<0, true>; // no Stride<> argument: the feature compiles away
};
template<typename A, typename... Rest>
struct FirstStride<A, Rest...> {
    using type = std::conditional_t<requires { A::kStride; }, A, typename FirstStride<Rest...>::type>;
};
} // namespace detail

// ---- type-erased view the graph / scheduler use (reference: BlockModel, BlockModel.hpp:334-574) -------------------------
class BlockModel {
public:
    virtual ~BlockModel()                                                     = default;
    virtual work::Result     work(std::size_t requested)                      = 0;
    virtual void             init()                                           = 0;
    virtual std::string_view name() const                                     = 0;
    virtual std::string_view typeName() const                                 = 0;
    virtual ComputeDomain    domain() const                                   = 0;
    virtual bool             runsOnDevice() const                             = 0;
    virtual std::size_t      inputCount() const                               = 0;
    virtual std::size_t      outputCount() const                              = 0;
    virtual std::size_t      inputItemBytes(std::size_t index) const          = 0;
    virtual std::size_t      outputItemBytes(std::size_t index) const         = 0;
    virtual int              inputPortIndex(std::string_view portName) const  = 0;
    virtual int              outputPortIndex(std::string_view portName) const = 0;
    virtual void             bindInput(std::size_t index, std::shared_ptr<EdgeBuffer> edge, int reader = 0) = 0;
    virtual void             bindOutput(std::size_t index, std::shared_ptr<EdgeBuffer> edge) = 0;
    virtual bool             inputOnDevice(std::size_t index) const           = 0; // which memory the port wants its edge in
    virtual bool             outputOnDevice(std::size_t index) const          = 0;
    virtual int              inputDevice(std::size_t index) const             = 0; // CUDA device of a device-side port's edge
    virtual int              outputDevice(std::size_t index) const            = 0;
    virtual int              workDevice() const                               = 0; // device whose stream runs this block; -1: host only
    virtual void             setStream(void* stream)                          = 0;
    virtual int              streamRole() const                               = 0; // 0 compute, 1 host->device copies, 2 device->host copies
    virtual void             start()                                          = 0; // optional user hook, once, before the first work()
    virtual std::size_t      inputHistoryItems(std::size_t index) const       = 0; // past items the block wants to find in front of its input spans
    virtual bool             chunksIndependent()                              = 0; // work chunks carry no state from one to the next
    virtual void             setStreams(std::vector<void*> streams)           = 0; // several streams: consecutive chunks rotate over them
    virtual std::size_t      inputChunkSize() const                           = 0; // after init(): the resampling ratio's two sides
    virtual std::size_t      outputChunkSize() const                          = 0;
    virtual property_map     settings()                                       = 0;
    virtual void*            raw()                                            = 0;
};

// ---- Block<Derived> --------------------------------------------------------------------------------------------------
template<typename Derived, typename... Arguments>
class Block {
public:
    using ResamplingControl = typename detail::FirstResampling<Arguments...>::type;
    using StrideControl     = typename detail::FirstStride<Arguments...>::type;

    // settings every block has (Block.hpp:708-713)
    std::string name;
    std::string compute_domain   = "host";
    std::size_t input_chunk_size  = ResamplingControl::kInputChunkSize;
    std::size_t output_chunk_size = ResamplingControl::kOutputChunkSize;
    std::size_t stride            = StrideControl::kStride; // Block.hpp:710 (active when != 0 and != input_chunk_size)

    Block() = default;
    explicit Block(property_map initialSettings) : _stagedSettings(std::move(initialSettings)) {}

    // staged like the reference (Block.hpp:449-451): applied at init() and at the top of the next work()
    void setSettings(property_map newSettings) {
        for (auto& [key, value] : newSettings) {
            _stagedSettings.insert_or_assign(key, std::move(value));
        }
    }

    [[nodiscard]] property_map currentSettings() {
        property_map all{{"name", name}, {"compute_domain", compute_domain}, {"input_chunk_size", static_cast<std::uint64_t>(input_chunk_size)}, {"output_chunk_size", static_cast<std::uint64_t>(output_chunk_size)}};
        forEachSetting([&](std::string_view key, auto& member) {
            using M = std::remove_cvref_t<decltype(member)>;
            if constexpr (is_annotated<M>::value) {
                all.insert_or_assign(std::string(key), valueOf(member.value));
            } else {
                all.insert_or_assign(std::string(key), valueOf(member));
            }
        });
        return all;
    }

    void init() {
        if (name.empty()) {
            name = std::string(Derived::gr_type_name());
        }
        applyStagedSettings();
        _domain = ComputeDomain::parse(compute_domain);
        if (_domain.isCuda() && !kHasCudaBody) {
            _warnedFallback = true; // reference behaviour (Block.hpp:1857-1862): warn once, run the host body
            _domain         = ComputeDomain{};
        }
        if (!_domain.isCuda() && !kHasHostBody) {
            throw exception(std::string(Derived::gr_type_name()) + ": compute_domain '" + compute_domain + "' is not a CUDA device and this block has no host implementation (no CPU fallback on the accelerated path)");
        }
    }

    [[nodiscard]] bool          runsOnDevice() const noexcept { return _domain.isCuda(); }
    [[nodiscard]] bool          warnedDeviceFallback() const noexcept { return _warnedFallback; }
    [[nodiscard]] ComputeDomain domain() const { return _domain; }
    void                        setStream(void* stream) noexcept { _stream = stream; }
    // A device block whose work chunks do not depend on each other (`bool chunksIndependent()` returns true: no state is
    // carried from chunk to chunk) may be given several streams: chunk k is issued on stream k mod n, so the tail of one
    // chunk's kernels overlaps the head of the next chunk's -- what keeps the SMs busy when chunks are small. The edges'
    // events carry every dependency between the streams.
    void setStreams(std::vector<void*> streams) {
        _streams = std::move(streams);
        if (!_streams.empty()) {
            _stream = _streams.front();
        }
    }
    [[nodiscard]] bool hasIndependentChunks() {
        if constexpr (requires(Derived& d) { d.chunksIndependent(); }) {
            return self().chunksIndependent();
        } else {
            return false;
        }
    }
    [[nodiscard]] void*         stream() const noexcept { return _stream; }

    work::Result work(std::size_t requested = std::numeric_limits<std::size_t>::max()) {
        if (_done) {
            return {requested, 0, work::Status::DONE};
        }
        if (!_stagedSettings.empty()) {
            applyStagedSettings();
        }
        return workInternal(requested);
    }

    void requestStop() noexcept { _stopRequested = true; }

    // A block with memory of its input (the FIR's past samples) may declare `std::size_t inputHistoryItems() const`: if its
    // input edge is an HBM ring it alone reads, the ring keeps that many past items in front of every span and the block
    // needs no state of its own (inputHistoryGranted() tells it whether the edge does)
    [[nodiscard]] std::size_t wantedInputHistory() const {
        if constexpr (requires(const Derived& d) { d.inputHistoryItems(); }) {
            return self().inputHistoryItems();
        } else {
            return 0;
        }
    }

    // the reference's optional lifecycle hook `void start()` (Block.hpp:598-607): the scheduler calls it once, after init()
    // and after the block got its stream, with the block's device current -- device blocks create their plans here so that
    // the first work chunk does not stall on allocations
    void invokeStart() {
        if constexpr (requires(Derived& d) { d.start(); }) {
            self().start();
        }
    }

protected:
    // a source that fills only part of the span it was handed publishes just that part (reference: OutputSpan::publish(n))
    void publishOnly(std::size_t nSamples) noexcept { _publishOverride = nSamples; }
    // tag on sample `offset` of the chunk being produced, on every output (reference: OutputSpan::publishTag)
    void publishTag(property_map map, std::size_t offset = 0) { _userTags.push_back(Tag{offset, std::move(map)}); }
    // the merged tag that arrived with the first sample of the chunk being processed (reference: Block::mergedInputTag)
    [[nodiscard]] const Tag& mergedInputTag() const noexcept { return _mergedInputTag; }
    [[nodiscard]] bool       inputTagsPresent() const noexcept { return !_mergedInputTag.map.empty(); }

    // waits for everything this block has queued so far (its stream, and every stream of a rotation): what a block does
    // before it replaces device data that queued launches still read (a settings change is rare)
    void synchronizeStreams() {
        if (_stream != nullptr) {
            gr4b200_stream_synchronize(_stream);
        }
        for (void* s : _streams) {
            if (s != _stream) {
                gr4b200_stream_synchronize(s);
            }
        }
    }

    // items of stream history in front of every input span (0: the block keeps its own state)
    [[nodiscard]] std::size_t inputHistoryGranted() {
        std::size_t granted = 0;
        forEachPort<PortDirection::INPUT>([&](std::size_t, std::string_view, auto& port) {
            if (port.edge) {
                granted = port.edge->historyItems();
            }
        });
        return granted;
    }

    Derived&       self() noexcept { return *static_cast<Derived*>(this); }
    const Derived& self() const noexcept { return *static_cast<const Derived*>(this); }

private:
    template<typename F>
    void forEachMember(F&& f) {
        auto           members = self().gr_members();
        constexpr auto names   = Derived::gr_member_names();
        [&]<std::size_t... I>(std::index_sequence<I...>) { (f(names[I], std::get<I>(members)), ...); }(std::make_index_sequence<std::tuple_size_v<decltype(members)>>{});
    }
    template<typename F>
    void forEachSetting(F&& f) {
        forEachMember([&](std::string_view key, auto& member) {
            if constexpr (!is_port<std::remove_cvref_t<decltype(member)>>::value && !is_port_vector<std::remove_cvref_t<decltype(member)>>::value) {
                f(key, member);
            }
        });
    }

    void applyStagedSettings() {
        property_map oldSettings, applied;
        for (auto& [key, value] : _stagedSettings) {
            bool known = false;
            if (key == "name") {
                known = assignFromValue(name, value);
            } else if (key == "compute_domain") {
                known = assignFromValue(compute_domain, value);
            } else if (key == "input_chunk_size") {
                known = assignFromValue(input_chunk_size, value);
            } else if (key == "output_chunk_size") {
                known = assignFromValue(output_chunk_size, value);
            } else if (key == "stride") {
                known = assignFromValue(stride, value);
            }
            forEachSetting([&](std::string_view memberName, auto& member) {
                if (memberName != key) {
                    return;
                }
                using M = std::remove_cvref_t<decltype(member)>;
                if constexpr (std::is_const_v<std::remove_reference_t<decltype(member)>>) {
                    return;
                } else if constexpr (is_annotated<M>::value) {
                    oldSettings.insert_or_assign(key, valueOf(member.value));
                    known = assignFromValue(member.value, value);
                } else {
                    oldSettings.insert_or_assign(key, valueOf(member));
                    known = assignFromValue(member, value);
                }
            });
            if (known) {
                applied.insert_or_assign(key, value);
                if (tag::isDefaultTag(key)) { // changed stream-related settings travel downstream as a tag (Block.hpp:1103-1111)
                    _pendingForward.insert_or_assign(key, value);
                }
            }
        }
        _stagedSettings.clear();
        if constexpr (requires(Derived& d, const property_map& m) { d.settingsChanged(m, m); }) {
            self().settingsChanged(oldSettings, applied); // also on the first application, like the reference's init()
        }
    }

    // ---- port plumbing ------------------------------------------------------------------------------------------------
    template<PortDirection Dir, typename F>
    void forEachPort(F&& f) {
        std::size_t index = 0;
        forEachMember([&](std::string_view key, auto& member) {
            using M = std::remove_cvref_t<decltype(member)>;
            if constexpr (is_port<M>::value) {
                if constexpr (M::direction == Dir) {
                    f(index++, key, member);
                }
            } else if constexpr (is_port_vector<M>::value) {
                if constexpr (M::value_type::direction == Dir) {
                    for (std::size_t k = 0; k < member.size(); ++k) {
                        const std::string indexed = std::string(key) + "#" + std::to_string(k);
                        f(index++, std::string_view(indexed), member[k]);
                    }
                }
            }
        });
    }

public:
    // used by BlockWrapper
    // port collections are sized by settings (`n_inputs`, `n_outputs`): they must exist when Graph::connect looks a port up,
    // which happens before init() applies the staged settings (the reference applies them inside emplaceBlock)
    void preparePorts() {
        if constexpr (requires(Derived& d) { d.in.resize(std::size_t{}); d.n_inputs; }) {
            if (const auto it = _stagedSettings.find("n_inputs"); it != _stagedSettings.end() && it->second.holdsNumber()) {
                self().in.resize(static_cast<std::size_t>(it->second.asDouble()));
            }
        }
        if constexpr (requires(Derived& d) { d.out.resize(std::size_t{}); d.n_outputs; }) {
            if (const auto it = _stagedSettings.find("n_outputs"); it != _stagedSettings.end() && it->second.holdsNumber()) {
                self().out.resize(static_cast<std::size_t>(it->second.asDouble()));
            }
        }
    }
    std::size_t portCount(PortDirection dir) {
        std::size_t n = 0;
        if (dir == PortDirection::INPUT) {
            forEachPort<PortDirection::INPUT>([&](std::size_t, std::string_view, auto&) { ++n; });
        } else {
            forEachPort<PortDirection::OUTPUT>([&](std::size_t, std::string_view, auto&) { ++n; });
        }
        return n;
    }
    std::size_t portItemBytes(PortDirection dir, std::size_t index) {
        std::size_t bytes = 0;
        auto        probe = [&](std::size_t i, std::string_view, auto& port) {
            if (i == index) {
                bytes = sizeof(typename std::remove_cvref_t<decltype(port)>::value_type);
            }
        };
        dir == PortDirection::INPUT ? forEachPort<PortDirection::INPUT>(probe) : forEachPort<PortDirection::OUTPUT>(probe);
        return bytes;
    }
    int portIndex(PortDirection dir, std::string_view portName) {
        int  found = -1;
        auto probe = [&](std::size_t i, std::string_view key, auto&) {
            if (key == portName) {
                found = static_cast<int>(i);
            }
        };
        dir == PortDirection::INPUT ? forEachPort<PortDirection::INPUT>(probe) : forEachPort<PortDirection::OUTPUT>(probe);
        return found;
    }
    void bindPort(PortDirection dir, std::size_t index, std::shared_ptr<EdgeBuffer> edge, int reader = 0) {
        auto bind = [&](std::size_t i, std::string_view, auto& port) {
            if (i == index) {
                port.edge   = edge;
                port.reader = reader;
            }
        };
        dir == PortDirection::INPUT ? forEachPort<PortDirection::INPUT>(bind) : forEachPort<PortDirection::OUTPUT>(bind);
    }
    // which memory a port's edge lives in: device blocks take device edges; a block may override per side through
    // `static constexpr bool kInputOnDevice / kOutputOnDevice` (the explicit domain-crossing blocks H2D / D2H do)
    bool portOnDevice(PortDirection dir) const {
        if constexpr (requires { Derived::kInputOnDevice; Derived::kOutputOnDevice; }) {
            return dir == PortDirection::INPUT ? Derived::kInputOnDevice : Derived::kOutputOnDevice;
        } else {
            return runsOnDevice();
        }
    }
    // which CUDA device a device-side port's edge lives on: the block's own device, unless the block bridges devices
    // (`int inputCudaDevice() const / outputCudaDevice() const`: H2D, D2H, PeerCopy)
    int portDevice(PortDirection dir) const {
        if constexpr (requires(const Derived& d) { d.inputCudaDevice(); d.outputCudaDevice(); }) {
            return dir == PortDirection::INPUT ? self().inputCudaDevice() : self().outputCudaDevice();
        } else {
            return _domain.isCuda() ? _domain.cudaDevice() : 0;
        }
    }
    // which of a device's streams carries this block's work: kernels share the compute stream; the copy blocks declare
    // `static constexpr int kStreamRole = 1 (host -> device) / 2 (device -> host)` so that transfers in both directions
    // and kernels overlap, ordered only by the edges' events
    int streamRole() const {
        if constexpr (requires { Derived::kStreamRole; }) {
            return Derived::kStreamRole;
        } else {
            return 0;
        }
    }
    // the device whose stream this block's work is issued on; -1 for blocks that never touch a device
    int workDevice() const {
        if constexpr (requires(const Derived& d) { d.cudaDeviceForWork(); }) {
            return self().cudaDeviceForWork();
        } else {
            return _domain.isCuda() ? _domain.cudaDevice() : -1;
        }
    }

private:
    static constexpr bool kHasCudaBody = requires { &Derived::processBulk_cuda; } || requires { &Derived::template processBulk_cuda<void>; };
    static constexpr bool kHasHostBody = requires { &Derived::processBulk; } || requires { &Derived::processOne; } || requires { &Derived::template processOne<int>; };

    template<typename PortT>
    static auto* inputPointer(PortT& port, std::size_t n, void* stream) { return static_cast<const typename PortT::value_type*>(port.edge->getLinear(n, stream, port.reader)); }
    template<typename PortT>
    static auto* outputPointer(PortT& port, std::size_t n, void* stream) { return static_cast<typename PortT::value_type*>(port.edge->reserve(n, stream)); }

    work::Result workInternal(std::size_t requested) {
        // 1. how much can move: min over input edges of what is published, min over output edges of what is free
        // Stride<> (Block.hpp:1546-1574): one chunk of input_chunk_size per call, the inputs advance by `stride` instead
        bool strideActive = false;
        if constexpr (StrideControl::kEnabled) {
            strideActive = stride != 0 && stride != input_chunk_size;
        }
        std::size_t nAvailable = std::numeric_limits<std::size_t>::max(), nRoom = std::numeric_limits<std::size_t>::max();
        std::size_t nInputs = 0, nOutputs = 0;
        bool        upstreamDone = true, unconnected = false;
        forEachPort<PortDirection::INPUT>([&](std::size_t, std::string_view, auto& port) {
            ++nInputs;
            if (!port.edge) {
                unconnected = true;
                return;
            }
            // "the producer is done and nothing is in flight" is sampled BEFORE the item count: a copy that completes between
            // the two polls must not leave a stale count next to a fresh "done" (the block would take the edge for drained
            // and finish with the last span unread)
            upstreamDone = upstreamDone && port.edge->producerDone && !port.edge->publishPending(); // (spans still being copied will show up)
            nAvailable   = std::min({nAvailable, strideActive ? port.edge->pending(port.reader) : port.edge->available(port.reader), port.max_samples});
        });
        forEachPort<PortDirection::OUTPUT>([&](std::size_t, std::string_view, auto& port) {
            ++nOutputs;
            if (!port.edge) {
                unconnected = true;
                return;
            }
            nRoom = std::min({nRoom, port.edge->writable(), port.max_samples});
        });
        if (unconnected) {
            return {requested, 0, work::Status::ERROR};
        }
        if (strideActive && _strideCounter > 0 && nInputs > 0) { // samples still to be skipped in front of the next chunk
            const std::size_t toSkip = std::min(_strideCounter, nAvailable);
            if (toSkip > 0) {
                forEachPort<PortDirection::INPUT>([&](std::size_t, std::string_view, auto& port) { port.edge->consume(toSkip, _stream, port.reader); });
                _strideCounter -= toSkip;
                nAvailable -= toSkip;
            }
            if (_strideCounter > 0) {
                return upstreamDone ? finish(requested) : work::Result{requested, toSkip, work::Status::INSUFFICIENT_INPUT_ITEMS};
            }
        }
        // 1b. tags: the ones on the first available sample belong to this chunk; a later one ends the chunk in front of it
        _mergedInputTag = Tag{};
        const std::size_t inChunkForTags = std::max<std::size_t>(input_chunk_size, 1);
        forEachPort<PortDirection::INPUT>([&](std::size_t, std::string_view, auto& port) {
            const std::lock_guard<std::recursive_mutex> tagLock(port.edge->mutex());
            const std::size_t base = port.edge->itemsConsumed(port.reader);
            for (const Tag& t : port.edge->tags) {
                if (t.index < base) {
                    continue; // already delivered to this reader; kept for a slower reader of the same edge
                }
                if (t.index == base) {
                    for (const auto& [key, value] : t.map) {
                        _mergedInputTag.map.insert_or_assign(key, value);
                    }
                } else if (t.index < base + nAvailable) {
                    const std::size_t upTo = (t.index - base) / inChunkForTags * inChunkForTags;
                    if (upTo > 0) { // (a tag inside the very first chunk of a resampling block moves to the chunk start)
                        nAvailable = upTo;
                    } else {
                        for (const auto& [key, value] : t.map) {
                            _mergedInputTag.map.insert_or_assign(key, value);
                        }
                        continue;
                    }
                    break;
                } else {
                    break;
                }
            }
        });
        if (!_mergedInputTag.map.empty() && !_tagsApplied) { // settings auto-update from tags, then settingsChanged
            _tagsApplied = true;
            property_map       updates;
            const property_map current = currentSettings();
            for (const auto& [key, value] : _mergedInputTag.map) {
                const std::string_view bare = tag::settingsKey(key);
                forEachSetting([&](std::string_view memberName, auto&) {
                    const auto it = current.find(bare);
                    if (memberName == bare && (it == current.end() || !(it->second == value))) { // unchanged values do not re-trigger settingsChanged
                        updates.insert_or_assign(std::string(bare), value);
                    }
                });
            }
            if (!updates.empty()) {
                setSettings(std::move(updates));
                applyStagedSettings();
            }
        }
        // 2. whole chunks only (Block.hpp:1610-1635)
        const std::size_t inChunk = std::max<std::size_t>(input_chunk_size, 1), outChunk = std::max<std::size_t>(output_chunk_size, 1);
        std::size_t       chunks = std::numeric_limits<std::size_t>::max();
        if (nInputs > 0) {
            chunks = std::min(chunks, std::min(nAvailable, requested) / inChunk);
        }
        if (nOutputs > 0) {
            chunks = std::min(chunks, (nInputs == 0 ? std::min(nRoom, requested) : nRoom) / outChunk);
        }
        if (strideActive && chunks > 1) {
            chunks = 1; // with a stride only one chunk at a time (Block.hpp:1617-1620)
        }
        if (_stopRequested) {
            chunks = 0;
        }
        if (chunks == 0) {
            const bool drained = nInputs > 0 && upstreamDone && nAvailable < inChunk;
            if (drained || _stopRequested) {
                return finish(requested);
            }
            return {requested, 0, nInputs > 0 && nAvailable < inChunk ? work::Status::INSUFFICIENT_INPUT_ITEMS : work::Status::INSUFFICIENT_OUTPUT_ITEMS};
        }
        const std::size_t nIn = nInputs > 0 ? chunks * inChunk : 0, nOut = nOutputs > 0 ? chunks * outChunk : 0;

        // 3. run the user body on the spans (independent chunks rotate over the block's streams)
        if (_streams.size() > 1 && hasIndependentChunks()) { // a block that starts to carry state (a settings change) stays on one stream
            _stream = _streams[_rotation++ % _streams.size()];
        }
        work::Status status = dispatch(nIn, nOut);
        if (status == work::Status::ERROR) {
            return {requested, 0, status};
        }
        // 4. the whole chunk is consumed and published (Block.hpp:1329-1362); tags first, they sit on the chunk's first sample
        std::size_t nConsume = nIn;
        if (strideActive && nInputs > 0) { // advance by the stride: less than the chunk = overlap, more = the rest is skipped before the next chunk
            nConsume       = std::min(stride, nIn);
            _strideCounter = stride - nConsume;
        }
        forEachPort<PortDirection::INPUT>([&](std::size_t, std::string_view, auto& port) { port.edge->consume(nConsume, _stream, port.reader); });
        _tagsApplied               = false;
        const std::size_t nPublish = std::min(nOut, _publishOverride);
        _publishOverride           = std::numeric_limits<std::size_t>::max();
        if (nOutputs > 0) {
            property_map forwarded = outputTagMap();
            forEachPort<PortDirection::OUTPUT>([&](std::size_t, std::string_view, auto& port) {
                port.edge->publishTag(forwarded, 0);
                for (const Tag& t : _userTags) {
                    port.edge->publishTag(t.map, std::min(t.index, nPublish > 0 ? nPublish - 1 : 0));
                }
            });
        }
        _userTags.clear();
        _pendingForward.clear();
        forEachPort<PortDirection::OUTPUT>([&](std::size_t, std::string_view, auto& port) { port.edge->publish(nPublish, _stream); });
        bool edgeFailed = false;
        forEachPort<PortDirection::INPUT>([&](std::size_t, std::string_view, auto& port) { edgeFailed = edgeFailed || port.edge->failed; });
        forEachPort<PortDirection::OUTPUT>([&](std::size_t, std::string_view, auto& port) { edgeFailed = edgeFailed || port.edge->failed; });
        if (edgeFailed) {
            return {requested, 0, work::Status::ERROR};
        }
        if (status == work::Status::DONE) {
            finish(requested);
        }
        return {requested, nInputs > 0 ? nIn : nOut, status};
    }

    // what goes out with the first sample of this chunk: the merged input tag (keys that name one of this block's
    // settings carry the block's own, possibly just updated, value) plus this block's changed stream settings;
    // `sample_rate` is multiplied by output_chunk_size / input_chunk_size on resampling blocks
    property_map outputTagMap() {
        property_map out;
        auto         insert = [&](const std::string& wireKey, const Value& value) {
            const std::string_view bare = tag::settingsKey(wireKey);
            if (bare == tag::SAMPLE_RATE && input_chunk_size != output_chunk_size && input_chunk_size != 0 && value.holdsNumber()) {
                const float ratio = static_cast<float>(output_chunk_size) / static_cast<float>(input_chunk_size);
                out.insert_or_assign(wireKey, Value(ratio * static_cast<float>(value.asDouble())));
            } else {
                out.insert_or_assign(wireKey, value);
            }
        };
        const property_map own = _mergedInputTag.map.empty() ? property_map{} : currentSettings();
        for (const auto& [key, value] : _mergedInputTag.map) {
            const auto it = own.find(tag::settingsKey(key));
            insert(key, it != own.end() && it->first != "name" && it->first != "compute_domain" ? it->second : value);
        }
        for (const auto& [key, value] : _pendingForward) {
            const bool viaPrefixed = out.contains(std::string(tag::kPrefix) + key);
            if (!out.contains(key) && !viaPrefixed) {
                insert(key, value);
            }
        }
        return out;
    }

    work::Result finish(std::size_t requested) {
        _done = true;
        forEachPort<PortDirection::OUTPUT>([&](std::size_t, std::string_view, auto& port) {
            if (port.edge) {
                port.edge->producerDone = true;
            }
        });
        return {requested, 0, work::Status::DONE};
    }

    // single-input / single-output / source / sink bodies (the shapes on this path)
    work::Status dispatch(std::size_t nIn, std::size_t nOut) {
        auto members = self().gr_members();
        return std::apply([&](auto&... m) { return dispatchPorts(nIn, nOut, m...); }, members);
    }

    template<typename... Members>
    work::Status dispatchPorts(std::size_t nIn, std::size_t nOut, Members&... members) {
        // collect pointers to the (at most one) input and output port among the reflected members
        using InPortT  = FirstPort<PortDirection::INPUT, std::remove_cvref_t<Members>...>;
        using OutPortT = FirstPort<PortDirection::OUTPUT, std::remove_cvref_t<Members>...>;
        using InVecT   = FirstPortVector<PortDirection::INPUT, std::remove_cvref_t<Members>...>;
        if constexpr (!std::is_void_v<typename InVecT::type> && !std::is_void_v<typename OutPortT::type>) {
            // N inputs of one type, one output (MathOpMultiPortImpl, Math.hpp:73-108): all inputs advance together
            auto& ins  = pick<typename InVecT::type>(members...);
            auto& out  = pick<typename OutPortT::type>(members...);
            using TIn  = typename InVecT::type::value_type::value_type;
            using TOut = typename OutPortT::type::value_type;
            std::vector<const TIn*> sources(ins.size());
            for (std::size_t k = 0; k < ins.size(); ++k) {
                sources[k] = inputPointer(ins[k], nIn, _stream);
                if (sources[k] == nullptr) {
                    return work::Status::ERROR;
                }
            }
            TOut* dst = outputPointer(out, nOut, _stream);
            if (dst == nullptr) {
                return work::Status::ERROR;
            }
            if constexpr (requires { self().processBulk_cuda(_stream, sources.data(), sources.size(), dst, nIn); }) {
                if (runsOnDevice()) {
                    return self().processBulk_cuda(_stream, sources.data(), sources.size(), dst, nIn);
                }
            }
            if constexpr (requires(std::span<const std::span<const TIn>> v) { self().processBulk(v, std::span<TOut>{}); }) {
                std::vector<std::span<const TIn>> spans;
                for (const TIn* src : sources) {
                    spans.emplace_back(src, nIn);
                }
                return self().processBulk(std::span<const std::span<const TIn>>(spans), std::span<TOut>(dst, nOut));
            } else {
                return work::Status::ERROR;
            }
        } else if constexpr (!std::is_void_v<typename InPortT::type> && !std::is_void_v<typename OutPortT::type>) {
            auto& in  = pick<typename InPortT::type>(members...);
            auto& out = pick<typename OutPortT::type>(members...);
            using TIn = typename InPortT::type::value_type;
            using TOut = typename OutPortT::type::value_type;
            const TIn* src = inputPointer(in, nIn, _stream);
            TOut*      dst = outputPointer(out, nOut, _stream);
            if (src == nullptr || dst == nullptr) {
                return work::Status::ERROR;
            }
            if constexpr (kHasCudaBody) {
                if (runsOnDevice()) {
                    return self().processBulk_cuda(_stream, src, dst, nIn, nOut);
                }
            }
            if constexpr (requires { self().processBulk(std::span<const TIn>{}, std::span<TOut>{}); }) {
                return self().processBulk(std::span<const TIn>(src, nIn), std::span<TOut>(dst, nOut));
            } else if constexpr (requires(const TIn& v) { self().processOne(v); }) {
                for (std::size_t i = 0; i < nIn; ++i) {
                    dst[i] = self().processOne(src[i]);
                }
                return work::Status::OK;
            } else {
                return work::Status::ERROR;
            }
        } else if constexpr (std::is_void_v<typename InPortT::type> && !std::is_void_v<typename OutPortT::type>) { // source
            auto& out  = pick<typename OutPortT::type>(members...);
            using TOut = typename OutPortT::type::value_type;
            TOut* dst  = outputPointer(out, nOut, _stream);
            if (dst == nullptr) {
                return work::Status::ERROR;
            }
            if constexpr (kHasCudaBody) {
                if (runsOnDevice()) {
                    return self().processBulk_cuda(_stream, static_cast<const TOut*>(nullptr), dst, 0, nOut);
                }
            }
            if constexpr (requires { self().processBulk(std::span<TOut>{}); }) {
                return self().processBulk(std::span<TOut>(dst, nOut));
            } else {
                for (std::size_t i = 0; i < nOut; ++i) {
                    dst[i] = self().processOne();
                }
                return work::Status::OK;
            }
        } else if constexpr (!std::is_void_v<typename InPortT::type>) { // sink
            auto& in  = pick<typename InPortT::type>(members...);
            using TIn = typename InPortT::type::value_type;
            const TIn* src = inputPointer(in, nIn, _stream);
            if (src == nullptr) {
                return work::Status::ERROR;
            }
            if constexpr (kHasCudaBody) {
                if (runsOnDevice()) {
                    return self().processBulk_cuda(_stream, src, static_cast<TIn*>(nullptr), nIn, 0);
                }
            }
            if constexpr (requires { self().processBulk(std::span<const TIn>{}); }) {
                return self().processBulk(std::span<const TIn>(src, nIn));
            } else {
                for (std::size_t i = 0; i < nIn; ++i) {
                    self().processOne(src[i]);
                }
                return work::Status::OK;
            }
        } else {
            return work::Status::ERROR;
        }
    }

    template<PortDirection Dir, typename... Ms>
    struct FirstPort {
        using type = void;
    };
    template<PortDirection Dir, typename M, typename... Rest>
    struct FirstPort<Dir, M, Rest...> {
        static constexpr bool match = [] {
            if constexpr (is_port<M>::value) {
                return M::direction == Dir;
            } else {
                return false;
            }
        }();
        using type = std::conditional_t<match, M, typename FirstPort<Dir, Rest...>::type>;
    };
    template<PortDirection Dir, typename... Ms>
    struct FirstPortVector {
        using type = void;
    };
    template<PortDirection Dir, typename M, typename... Rest>
    struct FirstPortVector<Dir, M, Rest...> {
        static constexpr bool match = [] {
            if constexpr (is_port_vector<M>::value) {
                return M::value_type::direction == Dir;
            } else {
                return false;
            }
        }();
        using type = std::conditional_t<match, M, typename FirstPortVector<Dir, Rest...>::type>;
    };
    template<typename Wanted, typename First, typename... Rest>
    static Wanted& pick(First& first, Rest&... rest) {
        if constexpr (std::is_same_v<std::remove_cvref_t<First>, Wanted>) {
            return first;
        } else {
            return pick<Wanted>(rest...);
        }
    }

    property_map  _stagedSettings;
    ComputeDomain _domain{};
    void*         _stream         = nullptr;
    std::vector<void*> _streams;      // set for blocks with independent chunks: rotation
    std::size_t        _rotation = 0;
    Tag              _mergedInputTag;
    std::vector<Tag> _userTags;       // publishTag() calls of the running chunk (index = offset inside the chunk)
    property_map     _pendingForward; // changed stream-related settings not yet sent downstream
    bool             _tagsApplied = false;
    bool          _done           = false;
    bool          _stopRequested  = false;
    bool          _warnedFallback = false;
    std::size_t   _publishOverride = std::numeric_limits<std::size_t>::max();
    std::size_t   _strideCounter   = 0; // leftover stride from previous calls (Block.hpp:715)
};

// ---- BlockWrapper<T>: owns a block, exposes BlockModel (reference: BlockModel.hpp:668) ----------------------------------
template<typename TBlock>
class BlockWrapper final : public BlockModel {
public:
    explicit BlockWrapper(property_map initial) : _block(std::move(initial)) {}
    TBlock&          block() noexcept { return _block; }
    work::Result     work(std::size_t requested) override { return _block.work(requested); }
    void             init() override { _block.init(); }
    std::string_view name() const override { return _block.name; }
    std::string_view typeName() const override { return TBlock::gr_type_name(); }
    ComputeDomain    domain() const override { return _block.domain(); }
    bool             runsOnDevice() const override { return _block.runsOnDevice(); }
    std::size_t      inputCount() const override { return const_cast<TBlock&>(_block).portCount(PortDirection::INPUT); }
    std::size_t      outputCount() const override { return const_cast<TBlock&>(_block).portCount(PortDirection::OUTPUT); }
    std::size_t      inputItemBytes(std::size_t i) const override { return const_cast<TBlock&>(_block).portItemBytes(PortDirection::INPUT, i); }
    std::size_t      outputItemBytes(std::size_t i) const override { return const_cast<TBlock&>(_block).portItemBytes(PortDirection::OUTPUT, i); }
    int              inputPortIndex(std::string_view n) const override { return prepared().portIndex(PortDirection::INPUT, n); }
    int              outputPortIndex(std::string_view n) const override { return prepared().portIndex(PortDirection::OUTPUT, n); }
    void             bindInput(std::size_t i, std::shared_ptr<EdgeBuffer> e, int reader = 0) override { _block.bindPort(PortDirection::INPUT, i, std::move(e), reader); }
    void             bindOutput(std::size_t i, std::shared_ptr<EdgeBuffer> e) override { _block.bindPort(PortDirection::OUTPUT, i, std::move(e)); }
    bool             inputOnDevice(std::size_t) const override { return _block.portOnDevice(PortDirection::INPUT); }
    bool             outputOnDevice(std::size_t) const override { return _block.portOnDevice(PortDirection::OUTPUT); }
    int              inputDevice(std::size_t) const override { return _block.portDevice(PortDirection::INPUT); }
    int              outputDevice(std::size_t) const override { return _block.portDevice(PortDirection::OUTPUT); }
    int              workDevice() const override { return _block.workDevice(); }
    void             setStream(void* stream) override { _block.setStream(stream); }
    int              streamRole() const override { return _block.streamRole(); }
    void             start() override { _block.invokeStart(); }
    std::size_t      inputHistoryItems(std::size_t) const override { return _block.wantedInputHistory(); }
    bool             chunksIndependent() override { return _block.hasIndependentChunks(); }
    void             setStreams(std::vector<void*> streams) override { _block.setStreams(std::move(streams)); }
    std::size_t      inputChunkSize() const override { return _block.input_chunk_size; }
    std::size_t      outputChunkSize() const override { return _block.output_chunk_size; }
    property_map     settings() override { return _block.currentSettings(); }
    void*            raw() override { return &_block; }

private:
    TBlock& prepared() const {
        auto& block = const_cast<TBlock&>(_block);
        block.preparePorts();
        return block;
    }

public:

private:
    TBlock _block;
};

} // namespace gr
