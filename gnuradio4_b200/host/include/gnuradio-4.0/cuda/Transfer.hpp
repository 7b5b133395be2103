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
    [[nodiscard]] int inputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int outputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int cudaDeviceForWork() const { return static_cast<int>(device); }
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
    [[nodiscard]] int inputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int outputCudaDevice() const { return static_cast<int>(device); }
    [[nodiscard]] int cudaDeviceForWork() const { return static_cast<int>(device); }
    gr::work::Status processBulk(std::span<const T> deviceInput, std::span<T> output) {
        if (gr4b200_copy_d2h(output.data(), deviceInput.data(), deviceInput.size_bytes(), this->stream()) != GR4B200_OK) {
            return gr::work::Status::ERROR;
        }
        // the host consumer reads the span as soon as it is published: wait for the copy
        return gr4b200_stream_synchronize(this->stream()) == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
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
