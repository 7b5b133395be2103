// gr4b200 host layer -- the source / sink fixtures the reference's tests and benchmarks are built from
// (blocks/testing/include/gnuradio-4.0/testing/NullSources.hpp:16-247, TagMonitors.hpp TagSource/TagSink value paths).
#pragma once

#include <vector>

#include "../Block.hpp"

namespace gr::testing {

template<typename T>
struct NullSource : gr::Block<NullSource<T>> { // produces zeros forever (bounded by the sink)
    using gr::Block<NullSource<T>>::Block;
    gr::PortOut<T> out;
    GR_MAKE_REFLECTABLE(NullSource, out);
    [[nodiscard]] constexpr T processOne() const noexcept { return T{}; }
};

template<typename T>
struct ConstantSource : gr::Block<ConstantSource<T>> {
    using gr::Block<ConstantSource<T>>::Block;
    gr::PortOut<T> out;
    T              default_value{};
    gr::Size_t     n_samples_max = 0; // 0: infinite
    GR_MAKE_REFLECTABLE(ConstantSource, out, default_value, n_samples_max);
    gr::Size_t _produced = 0;
    gr::work::Status processBulk(std::span<T> output) {
        std::size_t n = output.size();
        if (n_samples_max > 0) {
            n = std::min<std::size_t>(n, n_samples_max - _produced);
        }
        std::fill_n(output.begin(), n, default_value);
        this->publishOnly(n);
        _produced += static_cast<gr::Size_t>(n);
        return n_samples_max > 0 && _produced >= n_samples_max ? gr::work::Status::DONE : gr::work::Status::OK;
    }
};

template<typename T>
struct CountingSource : gr::Block<CountingSource<T>> { // 0, 1, 2, ...
    using gr::Block<CountingSource<T>>::Block;
    gr::PortOut<T> out;
    gr::Size_t     n_samples_max = 0;
    GR_MAKE_REFLECTABLE(CountingSource, out, n_samples_max);
    gr::Size_t _count = 0;
    gr::work::Status processBulk(std::span<T> output) {
        std::size_t n = output.size();
        if (n_samples_max > 0) {
            n = std::min<std::size_t>(n, n_samples_max - _count);
        }
        for (std::size_t i = 0; i < n; ++i) {
            output[i] = static_cast<T>(_count++);
        }
        this->publishOnly(n);
        return n_samples_max > 0 && _count >= n_samples_max ? gr::work::Status::DONE : gr::work::Status::OK;
    }
};

// TagSource's `values` path: replays a vector, then DONE
template<typename T>
struct VectorSource : gr::Block<VectorSource<T>> {
    using gr::Block<VectorSource<T>>::Block;
    gr::PortOut<T> out;
    std::vector<T> values;
    GR_MAKE_REFLECTABLE(VectorSource, out);
    std::size_t _position = 0;
    gr::work::Status processBulk(std::span<T> output) {
        const std::size_t n = std::min(output.size(), values.size() - _position);
        std::copy_n(values.begin() + static_cast<std::ptrdiff_t>(_position), n, output.begin());
        std::fill(output.begin() + static_cast<std::ptrdiff_t>(n), output.end(), T{});
        _position += n;
        this->publishOnly(n);
        return _position >= values.size() ? gr::work::Status::DONE : gr::work::Status::OK;
    }
};

template<typename T>
struct Copy : gr::Block<Copy<T>> {
    using gr::Block<Copy<T>>::Block;
    gr::PortIn<T>  in;
    gr::PortOut<T> out;
    GR_MAKE_REFLECTABLE(Copy, in, out);
    [[nodiscard]] constexpr T processOne(const T& v) const noexcept { return v; }
};

template<typename T>
struct NullSink : gr::Block<NullSink<T>> {
    using gr::Block<NullSink<T>>::Block;
    gr::PortIn<T> in;
    GR_MAKE_REFLECTABLE(NullSink, in);
    void processOne(const T&) const noexcept {}
};

template<typename T>
struct CountingSink : gr::Block<CountingSink<T>> { // stops the graph after n_samples_max samples (NullSources.hpp:200-247)
    using gr::Block<CountingSink<T>>::Block;
    gr::PortIn<T> in;
    gr::Size_t    n_samples_max = 0;
    GR_MAKE_REFLECTABLE(CountingSink, in, n_samples_max);
    std::size_t _count = 0;
    gr::work::Status processBulk(std::span<const T> input) {
        _count += input.size();
        return n_samples_max > 0 && _count >= n_samples_max ? gr::work::Status::DONE : gr::work::Status::OK;
    }
};

// TagSink's `_samples` path: keeps everything it receives
template<typename T>
struct VectorSink : gr::Block<VectorSink<T>> {
    using gr::Block<VectorSink<T>>::Block;
    gr::PortIn<T>  in;
    std::vector<T> _samples;
    GR_MAKE_REFLECTABLE(VectorSink, in);
    gr::work::Status processBulk(std::span<const T> input) {
        _samples.insert(_samples.end(), input.begin(), input.end());
        return gr::work::Status::OK;
    }
};

} // namespace gr::testing
