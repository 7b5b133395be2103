// gr4b200 host layer -- gr::blocks::math::{Add,Subtract,Multiply,Divide}Const for this path
// (reference: blocks/math/include/gnuradio-4.0/math/Math.hpp:30-66). Host body = the reference's processOne; device
// body for std::complex<float> = one C-ABI call.
#pragma once

#include <algorithm>
#include <complex>
#include <functional>
#include <vector>

#include "../Block.hpp"

namespace gr::blocks::math {

template<typename T, typename op>
struct MathOpImpl : gr::Block<MathOpImpl<T, op>> {
    using gr::Block<MathOpImpl<T, op>>::Block;
    gr::PortIn<T>  in{};
    gr::PortOut<T> out{};
    T              value = static_cast<T>(1);
    GR_MAKE_REFLECTABLE(MathOpImpl, in, out, value);

    [[nodiscard]] constexpr T processOne(const T& a) const noexcept { return op()(a, value); }
    [[nodiscard]] bool        chunksIndependent() const { return true; } // const + noexcept processOne: no state at all

    gr::work::Status processBulk_cuda(void* stream, const T* input, T* output, std::size_t nIn, std::size_t /*nOut*/)
    requires std::is_same_v<T, std::complex<float>>
    {
        constexpr int code = std::is_same_v<op, std::plus<T>> ? GR4B200_OP_ADD : std::is_same_v<op, std::minus<T>> ? GR4B200_OP_SUBTRACT : std::is_same_v<op, std::multiplies<T>> ? GR4B200_OP_MULTIPLY : GR4B200_OP_DIVIDE;
        const int     rc   = gr4b200_mathop_const_cf32(stream, code, reinterpret_cast<const float*>(input), reinterpret_cast<float*>(output), nIn, value.real(), value.imag());
        return rc == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
    }
};

template<typename T>
using AddConst = MathOpImpl<T, std::plus<T>>;
template<typename T>
using SubtractConst = MathOpImpl<T, std::minus<T>>;
template<typename T>
using MultiplyConst = MathOpImpl<T, std::multiplies<T>>;
template<typename T>
using DivideConst = MathOpImpl<T, std::divides<T>>;

// MathOpMultiPortImpl (Math.hpp:73-108): out = in#0 op in#1 op ... (left fold), `n_inputs` ports of one type
template<typename T, typename op>
struct MathOpMultiPortImpl : gr::Block<MathOpMultiPortImpl<T, op>> {
    using gr::Block<MathOpMultiPortImpl<T, op>>::Block;
    std::vector<gr::PortIn<T>> in{};
    gr::PortOut<T>             out{};
    gr::Size_t                 n_inputs = 0;
    GR_MAKE_REFLECTABLE(MathOpMultiPortImpl, in, out, n_inputs);

    void settingsChanged(const gr::property_map&, const gr::property_map& newSettings) {
        if (newSettings.contains("n_inputs")) {
            in.resize(n_inputs);
        }
    }

    gr::work::Status processBulk(std::span<const std::span<const T>> ins, std::span<T> output) const {
        if (ins.empty()) {
            return gr::work::Status::ERROR;
        }
        std::copy(ins[0].begin(), ins[0].end(), output.begin());
        for (std::size_t k = 1; k < ins.size(); ++k) {
            std::transform(output.begin(), output.end(), ins[k].begin(), output.begin(), op{});
        }
        return gr::work::Status::OK;
    }

    gr::work::Status processBulk_cuda(void* stream, const T* const* inputs, std::size_t nInputs, T* output, std::size_t n)
    requires std::is_same_v<T, std::complex<float>>
    {
        constexpr int code = std::is_same_v<op, std::plus<T>> ? GR4B200_OP_ADD : std::is_same_v<op, std::minus<T>> ? GR4B200_OP_SUBTRACT : std::is_same_v<op, std::multiplies<T>> ? GR4B200_OP_MULTIPLY : GR4B200_OP_DIVIDE;
        const int     rc   = gr4b200_mathop_multi_cf32(stream, code, reinterpret_cast<const float* const*>(inputs), nInputs, reinterpret_cast<float*>(output), n);
        return rc == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
    }
};

template<typename T>
using Add = MathOpMultiPortImpl<T, std::plus<T>>;
template<typename T>
using Subtract = MathOpMultiPortImpl<T, std::minus<T>>;
template<typename T>
using Multiply = MathOpMultiPortImpl<T, std::multiplies<T>>;
template<typename T>
using Divide = MathOpMultiPortImpl<T, std::divides<T>>;

} // namespace gr::blocks::math
