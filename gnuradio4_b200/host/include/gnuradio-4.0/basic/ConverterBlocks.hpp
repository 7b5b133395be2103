// gr4b200 host layer -- gr::blocks::type::converter::{InterleavedToComplex, ComplexToInterleaved} for this path
// (reference: blocks/basic/include/gnuradio-4.0/basic/ConverterBlocks.hpp:233-277): interleaved (re, im) items of
// float / int16 / int8 <-> std::complex<float>. Host body = the reference's element loop; device body = one C-ABI call.
#pragma once

#include <complex>
#include <cstdint>
#include <span>
#include <type_traits>

#include "../Block.hpp"

namespace gr::blocks::type::converter {

namespace detail {
template<typename R>
constexpr int itemTypeOf() {
    static_assert(std::is_same_v<R, float> || std::is_same_v<R, std::int16_t> || std::is_same_v<R, std::int8_t>, "interleaved items: float, int16_t or int8_t");
    return std::is_same_v<R, float> ? GR4B200_ITEM_F32 : (std::is_same_v<R, std::int16_t> ? GR4B200_ITEM_I16 : GR4B200_ITEM_I8);
}
} // namespace detail

// two interleaved items in, one complex sample out (Resampling<2, 1, true>, ConverterBlocks.hpp:260)
template<typename T, typename R>
requires std::is_same_v<R, std::complex<float>>
struct InterleavedToComplex : gr::Block<InterleavedToComplex<T, R>, gr::Resampling<2, 1, true>> {
    using gr::Block<InterleavedToComplex<T, R>, gr::Resampling<2, 1, true>>::Block;
    gr::PortIn<T>  interleaved;
    gr::PortOut<R> out;
    GR_MAKE_REFLECTABLE(InterleavedToComplex, interleaved, out);

    [[nodiscard]] gr::work::Status processBulk(std::span<const T> items, std::span<R> samples) const noexcept {
        const T* pair = items.data(); // (re, im), (re, im), ...
        for (R& sample : samples) {
            sample = R(static_cast<float>(pair[0]), static_cast<float>(pair[1]));
            pair += 2;
        }
        return gr::work::Status::OK;
    }

    gr::work::Status processBulk_cuda(void* stream, const T* input, R* output, std::size_t /*nIn*/, std::size_t nOut) {
        const int rc = gr4b200_interleaved_to_complex_cf32(stream, detail::itemTypeOf<T>(), input, reinterpret_cast<float*>(output), nOut);
        return rc == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
    }
};

// one complex sample in, two interleaved items out (Resampling<1, 2, true>, ConverterBlocks.hpp:237)
template<typename T, typename R>
requires std::is_same_v<T, std::complex<float>>
struct ComplexToInterleaved : gr::Block<ComplexToInterleaved<T, R>, gr::Resampling<1, 2, true>> {
    using gr::Block<ComplexToInterleaved<T, R>, gr::Resampling<1, 2, true>>::Block;
    gr::PortIn<T>  in;
    gr::PortOut<R> interleaved;
    GR_MAKE_REFLECTABLE(ComplexToInterleaved, in, interleaved);

    [[nodiscard]] gr::work::Status processBulk(std::span<const T> samples, std::span<R> items) const noexcept {
        R* pair = items.data(); // (re, im) per sample, each component cast to R (truncation toward zero for integers)
        for (const T& sample : samples) {
            pair[0] = static_cast<R>(sample.real());
            pair[1] = static_cast<R>(sample.imag());
            pair += 2;
        }
        return gr::work::Status::OK;
    }

    gr::work::Status processBulk_cuda(void* stream, const T* input, R* output, std::size_t nIn, std::size_t /*nOut*/) {
        const int rc = gr4b200_complex_to_interleaved_cf32(stream, detail::itemTypeOf<R>(), reinterpret_cast<const float*>(input), output, nIn);
        return rc == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
    }
};

} // namespace gr::blocks::type::converter
