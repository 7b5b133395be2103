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
    gr::work::Status processBulk(std::span<const T> deviceInput, std::span<T> output) {
        if (gr4b200_copy_d2h(output.data(), deviceInput.data(), deviceInput.size_bytes(), this->stream()) != GR4B200_OK) {
            return gr::work::Status::ERROR;
        }
        // the host consumer reads the span as soon as it is published: wait for the copy
        return gr4b200_stream_synchronize(this->stream()) == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
    }
};

} // namespace gr::cuda
