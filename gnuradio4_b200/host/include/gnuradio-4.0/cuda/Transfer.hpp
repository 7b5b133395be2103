// gr4b200 host layer -- explicit domain-crossing blocks. The reference requires transitions between port domains to be
// explicit conversion blocks (core/README.md "Ports"; PortDomain<"GPU"> tag, Port.hpp:172-185); these are they.
#pragma once

#include "../Block.hpp"

namespace gr::cuda {

template<typename T>
struct H2D : gr::Block<H2D<T>> { // host edge in, HBM edge out
    using gr::Block<H2D<T>>::Block;
    gr::PortIn<T>  in;
    gr::PortOut<T> out;
    gr::Size_t     device = 0;
    GR_MAKE_REFLECTABLE(H2D, in, out, device);
    static constexpr bool kInputOnDevice = false, kOutputOnDevice = true;
    static constexpr int  kStreamRole    = 1; // its own stream: uploads overlap kernels and downloads
    [[nodiscard]] int inputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int outputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int cudaDeviceForWork() const { return static_cast<int>(device); }
    // asynchronous: the host edge's read cursor follows the copy (EdgeBuffer::consume records an event behind it), the
    // HBM ring's write cursor is ordered by its own publish event
    gr::work::Status processBulk(std::span<const T> input, std::span<T> deviceOutput) {
        return gr4b200_copy_h2d(deviceOutput.data(), input.data(), input.size_bytes(), this->stream()) == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
    }
};

template<typename T>
struct D2H : gr::Block<D2H<T>> { // HBM edge in, host edge out
    using gr::Block<D2H<T>>::Block;
    gr::PortIn<T>  in;
    gr::PortOut<T> out;
    gr::Size_t     device = 0;
    GR_MAKE_REFLECTABLE(D2H, in, out, device);
    static constexpr bool kInputOnDevice = true, kOutputOnDevice = false;
    static constexpr int  kStreamRole    = 2; // its own stream: downloads overlap kernels and uploads
    [[nodiscard]] int inputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int outputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int cudaDeviceForWork() const { return static_cast<int>(device); }
    // asynchronous: the host edge publishes the span when the copy's event has completed (EdgeBuffer::publish), so the
    // host consumer never sees bytes that are still travelling and this block never waits for the GPU
    gr::work::Status processBulk(std::span<const T> deviceInput, std::span<T> output) {
        return gr4b200_copy_d2h(output.data(), deviceInput.data(), deviceInput.size_bytes(), this->stream()) == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
    }
};

// The graph boundary without a staging copy: a pinned host array streamed straight into the first HBM edge, and the
// last HBM edge streamed straight into a pinned host array (what an SDR driver that DMAs into pinned memory, or a file
// mapped into it, hands over). VectorSource -> H2D / D2H -> VectorSink do the same through a host edge, at the price of
// one memcpy per sample on the launcher thread.
template<typename T>
struct HostSource : gr::Block<HostSource<T>> {
    using gr::Block<HostSource<T>>::Block;
    gr::PortOut<T> out;
    gr::Size_t     device = 0;
    GR_MAKE_REFLECTABLE(HostSource, out, device);
    static constexpr bool kInputOnDevice = false, kOutputOnDevice = true;
    static constexpr int  kStreamRole    = 1;
    [[nodiscard]] int inputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int outputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int cudaDeviceForWork() const { return static_cast<int>(device); }
    // `data` must stay valid (and should be pinned: gr4b200_malloc_host) until runAndWait returns
    void setData(const T* data, std::size_t size) noexcept {
        _data     = data;
        _size     = size;
        _position = 0;
    }
    gr::work::Status processBulk(std::span<T> deviceOutput) {
        const std::size_t n = std::min(deviceOutput.size(), _size - _position);
        if (n > 0 && gr4b200_copy_h2d(deviceOutput.data(), _data + _position, n * sizeof(T), this->stream()) != GR4B200_OK) {
            return gr::work::Status::ERROR;
        }
        _position += n;
        this->publishOnly(n);
        return _position >= _size ? gr::work::Status::DONE : gr::work::Status::OK;
    }
    const T*    _data     = nullptr;
    std::size_t _size     = 0;
    std::size_t _position = 0;
};

template<typename T>
struct HostSink : gr::Block<HostSink<T>> {
    using gr::Block<HostSink<T>>::Block;
    gr::PortIn<T> in;
    gr::Size_t    device      = 0;
    gr::Size_t    copy_offset = 0; // bytes into every item ...
    gr::Size_t    copy_bytes  = 0; // ... and how many of them travel to the host (0: the whole item)
    GR_MAKE_REFLECTABLE(HostSink, in, device, copy_offset, copy_bytes);
    static constexpr bool kInputOnDevice = true, kOutputOnDevice = false;
    static constexpr int  kStreamRole    = 2;
    [[nodiscard]] int inputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int outputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int cudaDeviceForWork() const { return static_cast<int>(device); }
    // `data` receives up to `capacity` items (of bytesPerItem() bytes each, packed); it is complete when runAndWait has
    // returned (the scheduler synchronises every stream at the end of the run). With copy_bytes set only that slice of
    // every item crosses the link -- e.g. the magnitude plane of an FFT frame: the reference builds a whole DataSet per
    // transform whether or not anybody reads the other three signals (fft.hpp:173-250); here the rest stays in HBM.
    void setBuffer(void* data, std::size_t capacity) noexcept {
        _data     = static_cast<std::byte*>(data);
        _capacity = capacity;
        _position = 0;
    }
    [[nodiscard]] std::size_t bytesPerItem() const noexcept { return copy_bytes > 0 ? copy_bytes : sizeof(T); }
    [[nodiscard]] std::size_t itemsReceived() const noexcept { return _position; }
    gr::work::Status processBulk(std::span<const T> deviceInput) {
        const std::size_t n = std::min(deviceInput.size(), _capacity - _position);
        if (n > 0) {
            int rc;
            if (copy_bytes > 0) {
                if (copy_offset + copy_bytes > sizeof(T)) {
                    return gr::work::Status::ERROR;
                }
                rc = gr4b200_copy_d2h_2d(_data + _position * copy_bytes, copy_bytes, reinterpret_cast<const std::byte*>(deviceInput.data()) + copy_offset, sizeof(T), copy_bytes, n, this->stream());
            } else {
                rc = gr4b200_copy_d2h(_data + _position * sizeof(T), deviceInput.data(), n * sizeof(T), this->stream());
            }
            if (rc != GR4B200_OK) {
                return gr::work::Status::ERROR;
            }
        }
        _position += n;
        return _position >= _capacity ? gr::work::Status::DONE : gr::work::Status::OK;
    }
    std::byte*  _data     = nullptr;
    std::size_t _capacity = 0;
    std::size_t _position = 0;
};

// Device-resident ends for throughput measurements: a source that replays a capture already in HBM (it fills its ring
// during the first turn with device-to-device copies and only publishes afterwards -- the ring then holds the same
// samples turn after turn) and a sink that consumes without looking.
template<typename T>
struct DeviceReplaySource : gr::Block<DeviceReplaySource<T>> {
    using gr::Block<DeviceReplaySource<T>>::Block;
    gr::PortOut<T> out;
    gr::Size_t     device        = 0;
    gr::Size_t     n_samples_max = 0;
    GR_MAKE_REFLECTABLE(DeviceReplaySource, out, device, n_samples_max);
    static constexpr bool kInputOnDevice = true, kOutputOnDevice = true;
    [[nodiscard]] int inputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int outputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int cudaDeviceForWork() const { return static_cast<int>(device); }
    void setCapture(const T* deviceData, std::size_t size) noexcept { // `size` >= the capacity of the output edge
        _capture = deviceData;
        _size    = size;
    }
    gr::work::Status processBulk(std::span<T> deviceOutput) {
        const std::size_t n = std::min<std::size_t>(deviceOutput.size(), n_samples_max - _produced);
        if (_produced + n <= _size && n > 0) { // first ring turn: the span has not been filled yet
            if (gr4b200_copy_d2d(deviceOutput.data(), _capture + _produced, n * sizeof(T), this->stream()) != GR4B200_OK) {
                return gr::work::Status::ERROR;
            }
        }
        _produced += n;
        this->publishOnly(n);
        return _produced >= n_samples_max ? gr::work::Status::DONE : gr::work::Status::OK;
    }
    const T*    _capture  = nullptr;
    std::size_t _size     = 0;
    std::size_t _produced = 0;
};

template<typename T>
struct DeviceNullSink : gr::Block<DeviceNullSink<T>> {
    using gr::Block<DeviceNullSink<T>>::Block;
    gr::PortIn<T> in;
    gr::Size_t    device = 0;
    GR_MAKE_REFLECTABLE(DeviceNullSink, in, device);
    static constexpr bool kInputOnDevice = true, kOutputOnDevice = true;
    [[nodiscard]] int inputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int outputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int cudaDeviceForWork() const { return static_cast<int>(device); }
    std::size_t       _count = 0;
    gr::work::Status processBulk(std::span<const T> deviceInput) {
        _count += deviceInput.size();
        return gr::work::Status::OK;
    }
};

// HBM edge on one GPU in, HBM edge on another GPU out: the inter-GPU edge of the pipelined mode inside one process
// (cudaMemcpyPeerAsync over NVLink into the consumer's ring, issued on the destination device's stream; the
// multi-process flavour is gnuradio4_b200.multigpu.PipelinedChain with ncclSend / ncclRecv). The reference's analogue
// is the hand-off between the job lists of its multi-threaded scheduler (Scheduler.hpp:1944-1951).
template<typename T>
struct PeerCopy : gr::Block<PeerCopy<T>> {
    using gr::Block<PeerCopy<T>>::Block;
    gr::PortIn<T>  in;
    gr::PortOut<T> out;
    gr::Size_t     source_device = 0;
    gr::Size_t     device        = 1; // destination
    GR_MAKE_REFLECTABLE(PeerCopy, in, out, source_device, device);
    static constexpr bool kInputOnDevice = true, kOutputOnDevice = true;
    [[nodiscard]] int inputCudaDevice() const { return static_cast<int>(source_device); }
    [[nodiscard]] int outputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int cudaDeviceForWork() const { return static_cast<int>(device); }
    void settingsChanged(const gr::property_map&, const gr::property_map&) {
        gr4b200_peer_enable(static_cast<int>(device), static_cast<int>(source_device)); // direct access where the topology allows it;
        gr4b200_peer_enable(static_cast<int>(source_device), static_cast<int>(device)); // cudaMemcpyPeerAsync works either way
    }
    gr::work::Status processBulk(std::span<const T> sourceDeviceInput, std::span<T> destinationDeviceOutput) {
        const int rc = gr4b200_peer_copy(destinationDeviceOutput.data(), static_cast<int>(device), sourceDeviceInput.data(), static_cast<int>(source_device), sourceDeviceInput.size_bytes(), this->stream());
        return rc == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
    }
};

} // namespace gr::cuda
