// Runtime part of the C ABI: device binding, memory, streams/events, the HBM edge ring and peer copies.
#include <atomic>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace gr4b200 {
namespace {
thread_local std::string tlsLastError;
std::mutex               smCountMutex;
std::unordered_map<int, int> smCountCache;
} // namespace

std::atomic<unsigned long long>& launchCounter() {
    static std::atomic<unsigned long long> counter{0};
    return counter;
}
void        setLastError(const std::string& message) { tlsLastError = message; }
const char* lastError() { return tlsLastError.c_str(); }

int smCount() {
    int device = 0;
    if (cudaGetDevice(&device) != cudaSuccess) {
        return 148;
    }
    std::lock_guard<std::mutex> lock(smCountMutex);
    auto                        it = smCountCache.find(device);
    if (it != smCountCache.end()) {
        return it->second;
    }
    int count = 148;
    cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, device);
    smCountCache[device] = count;
    return count;
}
} // namespace gr4b200

using namespace gr4b200;

namespace {
// The last kDepth (cursor value, event) pairs recorded for one ring cursor. A stream that needs the cursor to have
// reached `need` waits on the OLDEST retained event whose cursor covers it -- not on the latest one: with a ring two
// chunks deep the producer of chunk k+1 only has to wait for the consumer of chunk k-1, so filling one half really
// overlaps draining the other (the reference's readers and writers overlap the same way through their sequence
// numbers, CircularBuffer.hpp:531-577,839-865).
// cudaEventRecord needs the event and the stream on the same device, while cudaStreamWaitEvent accepts an event of any
// device. A bridge block (gr::cuda::PeerCopy) consumes from a ring of one GPU on a stream of another: an event
// therefore lives on the device that is current when it is recorded (re-created if that ever changes).
struct CursorEvents {
    static constexpr int kDepth = 8;
    cudaEvent_t  event[kDepth]  = {};
    int          device[kDepth] = {};
    uint64_t     cursor[kDepth] = {};
    uint64_t     count          = 0;       // events recorded so far
    // Lazy mode: as long as every stream that asked for this cursor is the stream that moves it, stream order IS the
    // dependency and nothing is recorded at all (a chain of kernels on one compute stream pays no event per chunk).
    // The first foreign stream that has to wait switches the cursor to eager mode for good.
    bool         eager          = false;
    uint64_t     latest         = 0;       // cursor value after the last move
    cudaStream_t lastStream     = nullptr; // stream of the last move
    bool         moved          = false;

    int recordNow(uint64_t cursorAfter, cudaStream_t recordingStream) {
        const int slot    = static_cast<int>(count % kDepth);
        int       current = 0;
        GR4B200_CUDA_TRY(cudaGetDevice(&current));
        if (event[slot] == nullptr || device[slot] != current) {
            if (event[slot] != nullptr) {
                cudaEventDestroy(event[slot]);
                event[slot] = nullptr;
            }
            GR4B200_CUDA_TRY(cudaEventCreateWithFlags(&event[slot], cudaEventDisableTiming));
            device[slot] = current;
        }
        // "cursor >= c" must imply "every earlier move has completed too". Moves recorded on ONE stream complete in order by
        // themselves; a move recorded on another stream than its predecessor first waits for the predecessor's event
        // (a block whose work chunks rotate over several streams): the events then complete in cursor order again.
        if (count > 0 && recordingStream != eventStream) {
            GR4B200_CUDA_TRY(cudaStreamWaitEvent(recordingStream, event[(count - 1) % kDepth], 0));
        }
        GR4B200_CUDA_TRY(cudaEventRecord(event[slot], recordingStream));
        cursor[slot] = cursorAfter;
        eventStream  = recordingStream;
        ++count;
        return GR4B200_OK;
    }
    cudaStream_t eventStream = nullptr; // stream of the latest recorded event
    int record(uint64_t cursorAfter, cudaStream_t recordingStream) {
        if (moved && recordingStream != lastStream && !eager) {
            // the cursor itself starts to move from a second stream (a block whose work rotates over streams): an event
            // behind everything the first stream was given, then eager mode
            eager = true;
            int current = 0;
            GR4B200_CUDA_TRY(cudaGetDevice(&current));
            const int streamDevice = lastDevice >= 0 ? lastDevice : current;
            if (streamDevice != current) {
                GR4B200_CUDA_TRY(cudaSetDevice(streamDevice));
            }
            const int status = recordNow(latest, lastStream);
            if (streamDevice != current) {
                cudaSetDevice(current);
            }
            if (status != GR4B200_OK) {
                return status;
            }
        }
        latest     = cursorAfter;
        lastStream = recordingStream;
        moved      = true;
        noteDevice();
        return eager ? recordNow(cursorAfter, recordingStream) : GR4B200_OK;
    }
    // make `waitingStream` wait until the cursor has reached `need` (need == 0: nothing to wait for)
    int waitUntil(uint64_t need, cudaStream_t waitingStream) {
        if (need == 0) {
            return GR4B200_OK;
        }
        if (!eager) {
            if (!moved || waitingStream == lastStream) {
                return GR4B200_OK; // same stream: stream order is the dependency
            }
            // first foreign waiter: an event behind everything the moving stream has been given so far covers `need`
            // (need <= latest: the host cursor check has passed); the recording stream's device must be current for it
            eager = true;
            int current = 0;
            GR4B200_CUDA_TRY(cudaGetDevice(&current));
            const int streamDevice = lastDevice >= 0 ? lastDevice : current;
            if (streamDevice != current) {
                GR4B200_CUDA_TRY(cudaSetDevice(streamDevice));
            }
            const int status = recordNow(latest, lastStream);
            if (streamDevice != current) {
                cudaSetDevice(current);
            }
            if (status != GR4B200_OK) {
                return status;
            }
        }
        for (uint64_t i = count > kDepth ? count - kDepth : 0; i < count; ++i) {
            const int slot = static_cast<int>(i % kDepth);
            if (cursor[slot] >= need) {
                return checkCuda(cudaStreamWaitEvent(waitingStream, event[slot], 0), "cudaStreamWaitEvent(ring cursor)");
            }
        }
        return fail("ring: cursor event missing (host cursors and recorded events disagree)");
    }
    int  lastDevice = -1; // device that was current at the last move (the moving stream's device)
    void noteDevice() {
        int current = -1;
        if (cudaGetDevice(&current) == cudaSuccess) {
            lastDevice = current;
        }
    }
    void destroy() {
        for (auto& e : event) {
            if (e != nullptr) {
                cudaEventDestroy(e);
                e = nullptr;
            }
        }
    }
};
} // namespace

struct gr4b200_ring {
    int          device       = 0;
    char*        storage      = nullptr; // historyBytes + capacity bytes
    char*        base         = nullptr; // storage + historyBytes
    size_t       capacity     = 0;
    size_t       historyBytes = 0;
    uint64_t     written      = 0; // bytes published (monotonic)
    uint64_t     reserved     = 0; // bytes handed out by reserve (>= written)
    CursorEvents published;        // recorded by publish on the producer's stream
    cudaEvent_t  prefixEvent       = nullptr; // behind the reader's latest refresh of the history prefix
    int          prefixEventDevice = -1;
    cudaStream_t prefixStream      = nullptr;
    // one writer, N readers (CircularBuffer.hpp:476-477): every reader has its own cursor and its own "consumed" events;
    // space is free once the slowest reader has passed it. Reader 0 exists from creation.
    static constexpr int kMaxReaders = 8;
    int          nReaders              = 1;
    uint64_t     consumed[kMaxReaders] = {}; // bytes consumed per reader (monotonic)
    CursorEvents consumedEvents[kMaxReaders];
    uint64_t     slowest() const {
        uint64_t m = consumed[0];
        for (int r = 1; r < nReaders; ++r) {
            m = consumed[r] < m ? consumed[r] : m;
        }
        return m;
    }
};

// the cursor store of an inter-process edge: behind the producing kernel in stream order, visible to the peer
static __global__ void storeValueKernel(unsigned* target, unsigned value) {
    *reinterpret_cast<volatile unsigned*>(target) = value;
    __threadfence_system();
}

extern "C" {

int         gr4b200_abi_version(void) { return GR4B200_ABI_VERSION; }
const char* gr4b200_last_error(void) { return lastError(); }

int gr4b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int gr4b200_init(int device) {
    GR4B200_CUDA_TRY(cudaSetDevice(device));
    GR4B200_CUDA_TRY(cudaFree(nullptr)); // force context creation on this thread
    return GR4B200_OK;
}

int gr4b200_device_sm_count(int device) {
    int count = 0;
    if (cudaDeviceGetAttribute(&count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return count;
}

void* gr4b200_malloc(size_t bytes) {
    void* p = nullptr;
    if (checkCuda(cudaMalloc(&p, bytes == 0 ? 1 : bytes), "cudaMalloc") != GR4B200_OK) {
        return nullptr;
    }
    return p;
}
int   gr4b200_free(void* devicePtr) { return checkCuda(cudaFree(devicePtr), "cudaFree"); }
void* gr4b200_malloc_host(size_t bytes) {
    void* p = nullptr;
    if (checkCuda(cudaMallocHost(&p, bytes == 0 ? 1 : bytes), "cudaMallocHost") != GR4B200_OK) {
        return nullptr;
    }
    return p;
}
int gr4b200_free_host(void* hostPtr) { return checkCuda(cudaFreeHost(hostPtr), "cudaFreeHost"); }
int gr4b200_memset(void* devicePtr, int value, size_t bytes, void* stream) { return checkCuda(cudaMemsetAsync(devicePtr, value, bytes, asStream(stream)), "cudaMemsetAsync"); }
int gr4b200_copy_h2d(void* devicePtr, const void* hostPtr, size_t bytes, void* stream) { return checkCuda(cudaMemcpyAsync(devicePtr, hostPtr, bytes, cudaMemcpyHostToDevice, asStream(stream)), "cudaMemcpyAsync(H2D)"); }
int gr4b200_copy_d2h(void* hostPtr, const void* devicePtr, size_t bytes, void* stream) { return checkCuda(cudaMemcpyAsync(hostPtr, devicePtr, bytes, cudaMemcpyDeviceToHost, asStream(stream)), "cudaMemcpyAsync(D2H)"); }
int gr4b200_copy_d2h_2d(void* hostPtr, size_t dstPitch, const void* devicePtr, size_t srcPitch, size_t widthBytes, size_t height, void* stream) { return checkCuda(cudaMemcpy2DAsync(hostPtr, dstPitch, devicePtr, srcPitch, widthBytes, height, cudaMemcpyDeviceToHost, asStream(stream)), "cudaMemcpy2DAsync(D2H)"); }
unsigned long long gr4b200_launch_count(void) { return gr4b200::launchCounter().load(std::memory_order_relaxed); }
int gr4b200_copy_d2d(void* dst, const void* src, size_t bytes, void* stream) { return checkCuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, asStream(stream)), "cudaMemcpyAsync(D2D)"); }

void* gr4b200_stream_create(void) {
    cudaStream_t s = nullptr;
    if (checkCuda(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate") != GR4B200_OK) {
        return nullptr;
    }
    return s;
}
int   gr4b200_stream_destroy(void* stream) { return checkCuda(cudaStreamDestroy(asStream(stream)), "cudaStreamDestroy"); }
int   gr4b200_stream_synchronize(void* stream) { return checkCuda(cudaStreamSynchronize(asStream(stream)), "cudaStreamSynchronize"); }
void* gr4b200_event_create(void) {
    cudaEvent_t e = nullptr;
    if (checkCuda(cudaEventCreate(&e), "cudaEventCreate") != GR4B200_OK) {
        return nullptr;
    }
    return e;
}
int gr4b200_event_destroy(void* event) { return checkCuda(cudaEventDestroy(static_cast<cudaEvent_t>(event)), "cudaEventDestroy"); }
int gr4b200_event_record(void* event, void* stream) { return checkCuda(cudaEventRecord(static_cast<cudaEvent_t>(event), asStream(stream)), "cudaEventRecord"); }
int gr4b200_stream_wait_event(void* stream, void* event) { return checkCuda(cudaStreamWaitEvent(asStream(stream), static_cast<cudaEvent_t>(event), 0), "cudaStreamWaitEvent"); }
int gr4b200_event_query(void* event) {
    const cudaError_t err = cudaEventQuery(static_cast<cudaEvent_t>(event));
    if (err == cudaSuccess) {
        return 1;
    }
    if (err == cudaErrorNotReady) {
        cudaGetLastError();
        return 0;
    }
    return checkCuda(err, "cudaEventQuery");
}
int gr4b200_event_synchronize(void* event) { return checkCuda(cudaEventSynchronize(static_cast<cudaEvent_t>(event)), "cudaEventSynchronize"); }
int gr4b200_event_elapsed_ms(void* start, void* stop, float* ms) { return checkCuda(cudaEventElapsedTime(ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)), "cudaEventElapsedTime"); }

// ---- HBM edge ring ----------------------------------------------------------------------------------------------
gr4b200_ring* gr4b200_ring_create(int device, size_t capacityBytes, size_t historyBytes) {
    if (capacityBytes == 0 || historyBytes > capacityBytes) {
        fail("ring: capacity must be > 0 and >= history");
        return nullptr;
    }
    if (checkCuda(cudaSetDevice(device), "cudaSetDevice") != GR4B200_OK) {
        return nullptr;
    }
    auto* ring         = new gr4b200_ring;
    ring->device       = device;
    ring->capacity     = capacityBytes;
    ring->historyBytes = (historyBytes + 255) / 256 * 256; // keep the base 256-byte aligned
    void* p            = nullptr;
    if (checkCuda(cudaMalloc(&p, ring->historyBytes + capacityBytes), "cudaMalloc(ring)") != GR4B200_OK) {
        delete ring;
        return nullptr;
    }
    ring->storage = static_cast<char*>(p);
    ring->base    = ring->storage + ring->historyBytes;
    cudaMemset(ring->storage, 0, ring->historyBytes + capacityBytes); // x[<0] = 0, like a freshly constructed history
    return ring;
}

int gr4b200_ring_destroy(gr4b200_ring* ring) {
    if (ring == nullptr) {
        return GR4B200_OK;
    }
    ring->published.destroy();
    if (ring->prefixEvent != nullptr) {
        cudaEventDestroy(ring->prefixEvent);
    }
    for (auto& events : ring->consumedEvents) {
        events.destroy();
    }
    const int status = checkCuda(cudaFree(ring->storage), "cudaFree(ring)");
    delete ring;
    return status;
}

size_t gr4b200_ring_capacity(const gr4b200_ring* ring) { return ring->capacity; }

int gr4b200_ring_add_reader(gr4b200_ring* ring) {
    if (ring == nullptr) {
        return fail("ring_add_reader: null ring");
    }
    if (ring->historyBytes > 0) {
        return fail("ring_add_reader: a ring with history has one reader (the reader maintains the history prefix)");
    }
    if (ring->nReaders >= gr4b200_ring::kMaxReaders) {
        return fail("ring_add_reader: at most 8 readers per edge");
    }
    if (ring->written != 0) {
        return fail("ring_add_reader: readers join before the first publish (the reference wires all readers at connect time)");
    }
    const int reader = ring->nReaders;
    ring->nReaders   = reader + 1;
    return reader;
}

size_t gr4b200_ring_available_for(const gr4b200_ring* ring, int reader) {
    if (reader < 0 || reader >= ring->nReaders) {
        return 0;
    }
    const size_t pending    = static_cast<size_t>(ring->written - ring->consumed[reader]);
    const size_t contiguous = ring->capacity - static_cast<size_t>(ring->consumed[reader] % ring->capacity);
    return pending < contiguous ? pending : contiguous;
}
size_t gr4b200_ring_available(const gr4b200_ring* ring) { return gr4b200_ring_available_for(ring, 0); }

size_t gr4b200_ring_writable(const gr4b200_ring* ring) {
    // a ring with history keeps the `history` bytes behind the reader's cursor intact: they are not free yet
    const size_t inUse      = static_cast<size_t>(ring->reserved - ring->slowest()) + ring->historyBytes;
    const size_t freeBytes  = inUse < ring->capacity ? ring->capacity - inUse : 0;
    const size_t contiguous = ring->capacity - static_cast<size_t>(ring->reserved % ring->capacity);
    return freeBytes < contiguous ? freeBytes : contiguous;
}

void* gr4b200_ring_reserve(gr4b200_ring* ring, size_t bytes, void* stream) {
    if (ring->reserved != ring->written) {
        fail("ring: previous reservation not published");
        return nullptr;
    }
    if (bytes > gr4b200_ring_writable(ring)) {
        fail("ring: not enough contiguous free space", GR4B200_INSUFFICIENT_OUTPUT_ITEMS);
        return nullptr;
    }
    // the producer stream waits until every reader has left THIS span: cursor >= reserved + bytes - capacity
    const uint64_t end  = ring->reserved + bytes + ring->historyBytes; // (+ the history behind the reader's cursor)
    const uint64_t need = end > ring->capacity ? end - ring->capacity : 0;
    for (int r = 0; r < ring->nReaders; ++r) {
        if (ring->consumedEvents[r].waitUntil(need, asStream(stream)) != GR4B200_OK) {
            return nullptr;
        }
    }
    void* p = ring->base + ring->reserved % ring->capacity;
    ring->reserved += bytes;
    return p;
}

int gr4b200_ring_publish(gr4b200_ring* ring, size_t bytes, void* stream) {
    if (bytes > ring->reserved - ring->written) {
        return fail("ring: publishing more than reserved");
    }
    ring->written += bytes;
    ring->reserved = ring->written; // a short publish gives the rest of the reservation back
    return ring->published.record(ring->written, asStream(stream));
}

const void* gr4b200_ring_get_for(gr4b200_ring* ring, int reader, size_t bytes, void* stream) {
    if (reader < 0 || reader >= ring->nReaders) {
        fail("ring: no such reader");
        return nullptr;
    }
    if (bytes > gr4b200_ring_available_for(ring, reader)) {
        fail("ring: not enough contiguous published data", GR4B200_INSUFFICIENT_INPUT_ITEMS);
        return nullptr;
    }
    if (ring->published.waitUntil(ring->consumed[reader] + bytes, asStream(stream)) != GR4B200_OK) { // the publish that covers this span
        return nullptr;
    }
    if (ring->prefixEvent != nullptr && ring->prefixStream != asStream(stream) && ring->consumed[reader] % ring->capacity == 0) {
        // the span at offset 0 has its history in the prefix: the copy that refreshed it ran on another stream
        if (checkCuda(cudaStreamWaitEvent(asStream(stream), ring->prefixEvent, 0), "cudaStreamWaitEvent(ring prefix)") != GR4B200_OK) {
            return nullptr;
        }
    }
    return ring->base + ring->consumed[reader] % ring->capacity;
}
const void* gr4b200_ring_get(gr4b200_ring* ring, size_t bytes, void* stream) { return gr4b200_ring_get_for(ring, 0, bytes, stream); }

size_t gr4b200_ring_pending_for(const gr4b200_ring* ring, int reader) {
    if (reader < 0 || reader >= ring->nReaders) {
        return 0;
    }
    return static_cast<size_t>(ring->written - ring->consumed[reader]);
}

int gr4b200_ring_read_for(gr4b200_ring* ring, int reader, size_t bytes, void* dst, void* stream) {
    if (reader < 0 || reader >= ring->nReaders) {
        return fail("ring: no such reader");
    }
    if (bytes > gr4b200_ring_pending_for(ring, reader)) {
        return fail("ring: not enough published data", GR4B200_INSUFFICIENT_INPUT_ITEMS);
    }
    if (const int status = ring->published.waitUntil(ring->consumed[reader] + bytes, asStream(stream)); status != GR4B200_OK) {
        return status;
    }
    const size_t begin = static_cast<size_t>(ring->consumed[reader] % ring->capacity);
    const size_t first = bytes < ring->capacity - begin ? bytes : ring->capacity - begin;
    GR4B200_CUDA_TRY(cudaMemcpyAsync(dst, ring->base + begin, first, cudaMemcpyDeviceToDevice, asStream(stream)));
    if (first < bytes) {
        GR4B200_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(dst) + first, ring->base, bytes - first, cudaMemcpyDeviceToDevice, asStream(stream)));
    }
    return GR4B200_OK;
}

int gr4b200_ring_consume_for(gr4b200_ring* ring, int reader, size_t bytes, void* stream) {
    if (reader < 0 || reader >= ring->nReaders) {
        return fail("ring: no such reader");
    }
    if (bytes > ring->written - ring->consumed[reader]) {
        return fail("ring: consuming more than published");
    }
    const uint64_t consumedBefore = ring->consumed[reader];
    ring->consumed[reader] += bytes;
    // keep `history` bytes in front of offset 0 valid: when the reader leaves the end of the ring, the last bytes of the ring
    // are copied in front of the base ON THE READER'S STREAM -- after its work on this span, before the consume event, so
    // the writer cannot touch the tail until the copy has read it, and after the reader's own work on the span at offset 0
    // of this turn, which read the old prefix. (A history ring has exactly one reader: see gr4b200_ring_add_reader.)
    // A reader whose chunks rotate over several streams: the copy also waits for the reader's earlier chunks of this turn
    // (they ran on other streams and may still read the old prefix), and the next chunk at offset 0 waits for the copy.
    if (ring->historyBytes > 0 && bytes > 0 && ring->consumed[reader] % ring->capacity == 0) {
        if (ring->consumedEvents[reader].waitUntil(consumedBefore, asStream(stream)) != GR4B200_OK) {
            return GR4B200_ERROR;
        }
        GR4B200_CUDA_TRY(cudaMemcpyAsync(ring->base - ring->historyBytes, ring->base + ring->capacity - ring->historyBytes, ring->historyBytes, cudaMemcpyDeviceToDevice, asStream(stream)));
        int current = 0;
        GR4B200_CUDA_TRY(cudaGetDevice(&current));
        if (ring->prefixEvent == nullptr || ring->prefixEventDevice != current) {
            if (ring->prefixEvent != nullptr) {
                cudaEventDestroy(ring->prefixEvent);
            }
            GR4B200_CUDA_TRY(cudaEventCreateWithFlags(&ring->prefixEvent, cudaEventDisableTiming));
            ring->prefixEventDevice = current;
        }
        GR4B200_CUDA_TRY(cudaEventRecord(ring->prefixEvent, asStream(stream)));
        ring->prefixStream = asStream(stream);
    }
    return ring->consumedEvents[reader].record(ring->consumed[reader], asStream(stream));
}
int gr4b200_ring_consume(gr4b200_ring* ring, size_t bytes, void* stream) { return gr4b200_ring_consume_for(ring, 0, bytes, stream); }

// ---- inter-GPU edges ----------------------------------------------------------------------------------------------
int gr4b200_peer_enable(int device, int peerDevice) {
    if (device == peerDevice) {
        return GR4B200_OK;
    }
    int canAccess = 0;
    GR4B200_CUDA_TRY(cudaDeviceCanAccessPeer(&canAccess, device, peerDevice));
    if (!canAccess) {
        return fail("peer access not supported between these devices");
    }
    int previous = 0;
    GR4B200_CUDA_TRY(cudaGetDevice(&previous));
    GR4B200_CUDA_TRY(cudaSetDevice(device));
    const cudaError_t err = cudaDeviceEnablePeerAccess(peerDevice, 0);
    cudaSetDevice(previous);
    if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled) {
        return checkCuda(err, "cudaDeviceEnablePeerAccess");
    }
    cudaGetLastError();
    return GR4B200_OK;
}

int gr4b200_ipc_export(void* devicePtr, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the C ABI");
    cudaIpcMemHandle_t handle;
    GR4B200_CUDA_TRY(cudaIpcGetMemHandle(&handle, devicePtr));
    std::memcpy(handle64, &handle, sizeof handle);
    return GR4B200_OK;
}
void* gr4b200_ipc_open(const void* handle64) {
    cudaIpcMemHandle_t handle;
    std::memcpy(&handle, handle64, sizeof handle);
    void* mapped = nullptr;
    if (checkCuda(cudaIpcOpenMemHandle(&mapped, handle, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle") != GR4B200_OK) {
        return nullptr;
    }
    return mapped;
}
int gr4b200_ipc_close(void* mappedPtr) { return checkCuda(cudaIpcCloseMemHandle(mappedPtr), "cudaIpcCloseMemHandle"); }

int gr4b200_stream_write_value32(void* stream, void* devicePtr, unsigned value) {
    storeValueKernel<<<1, 1, 0, asStream(stream)>>>(static_cast<unsigned*>(devicePtr), value);
    return checkLaunch("storeValueKernel");
}
int gr4b200_stream_wait_value32(void* stream, void* devicePtr, unsigned value) {
    // driver entry point through the runtime: no link-time dependency on libcuda (the library also loads on hosts without a GPU)
    using WaitFn = int (*)(void*, unsigned long long, unsigned, unsigned);
    static WaitFn waitValue = [] {
        void*                            fn     = nullptr;
        cudaDriverEntryPointQueryResult status = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &status) != cudaSuccess || status != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            fn = nullptr;
        }
        return reinterpret_cast<WaitFn>(fn);
    }();
    if (waitValue == nullptr) {
        return fail("stream_wait_value32: the driver has no cuStreamWaitValue32");
    }
    constexpr unsigned kWaitGeq = 0x0; // CU_STREAM_WAIT_VALUE_GEQ
    const int          rc       = waitValue(stream, reinterpret_cast<unsigned long long>(devicePtr), value, kWaitGeq);
    return rc == 0 ? GR4B200_OK : fail("cuStreamWaitValue32 failed with CUresult " + std::to_string(rc));
}

int gr4b200_peer_copy(void* dst, int dstDevice, const void* src, int srcDevice, size_t bytes, void* stream) { return checkCuda(cudaMemcpyPeerAsync(dst, dstDevice, src, srcDevice, bytes, asStream(stream)), "cudaMemcpyPeerAsync"); }

} // extern "C"
