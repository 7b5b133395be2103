// gr4b200 host layer -- gr::filter::fir_filter, BasicDecimatingFilter (FIR), Decimator for this path
// (reference: blocks/filter/include/gnuradio-4.0/filter/time_domain_filter.hpp:20-48, :129-245). Device only.
// fir_filter here is instantiable for float (as registered in the reference) and std::complex<float> (the build target:
// the same real taps applied to re and im).
#pragma once

#include <algorithm>
#include <bit>
#include <complex>
#include <vector>

#include "../Block.hpp"

namespace gr::filter {

enum class FilterType { FIR, IIR };
enum class Type { LOWPASS = 0, HIGHPASS, BANDPASS, BANDSTOP };

namespace detail {
template<typename T>
inline int runFir(gr4b200_fir_plan* plan, void* stream, const T* input, T* output, std::size_t nIn, bool historyInStream) {
    if constexpr (std::is_same_v<T, float>) {
        return historyInStream ? gr4b200_fir_f32_contiguous(plan, stream, input, output, nIn) : gr4b200_fir_f32(plan, stream, input, output, nIn);
    } else {
        return historyInStream ? gr4b200_fir_cf32_contiguous(plan, stream, reinterpret_cast<const float*>(input), reinterpret_cast<float*>(output), nIn) : gr4b200_fir_cf32(plan, stream, reinterpret_cast<const float*>(input), reinterpret_cast<float*>(output), nIn);
    }
}
// the past samples a filter of nTaps coefficients reads, in the granularity the kernels stage them (16 samples)
inline std::size_t firHistoryItems(std::size_t nTaps) { return nTaps > 1 ? (nTaps - 1 + 15) / 16 * 16 : 0; }
// capacity of the reference's HistoryBuffer for a first `b` of nTaps coefficients (time_domain_filter.hpp:23, :40-42)
inline std::size_t firReferenceCapacity(std::size_t nTaps) { return nTaps > 32 ? std::bit_ceil(nTaps) : 32; }
} // namespace detail

template<typename T>
requires(std::is_same_v<T, float> || std::is_same_v<T, std::complex<float>>)
struct fir_filter : gr::Block<fir_filter<T>> {
    using gr::Block<fir_filter<T>>::Block;
    gr::PortIn<T>      in;
    gr::PortOut<T>     out;
    std::vector<float> b{1.f}; // feed-forward coefficients
    bool               exact = true; // reference summation order and rounding (bit-identical); false: fused multiply-add
    bool               overlap_save = false; // tolerance mode through the 4096-point transform (complex<float>): HBM bound
    GR_MAKE_REFLECTABLE(fir_filter, in, out, b, exact, overlap_save);

    [[nodiscard]] int planMode() const { return overlap_save ? GR4B200_FIR_OVERLAP_SAVE : (exact ? GR4B200_FIR_EXACT : GR4B200_FIR_FAST); }

    ~fir_filter() { gr4b200_fir_plan_destroy(_plan); }

    // New coefficients for a running filter (time_domain_filter.hpp:39-43): the reference replaces -- and thereby zeroes --
    // its HistoryBuffer only when `b` no longer fits it (capacity 32, then bit_ceil(b.size())); otherwise the past samples
    // stay and the next outputs are the new coefficients over the old samples. The plan does the same
    // (gr4b200_fir_plan_set_taps). When the past samples are read from the input ring instead of the plan's state, a `b`
    // that fits finds them there (the block asks the ring for the reference's whole capacity); one that does not fit must
    // see zeros, so the block moves to the plan's freshly zeroed state for good.
    void settingsChanged(const gr::property_map& /*oldSettings*/, const gr::property_map& newSettings) {
        const bool modeChanged = newSettings.contains("exact") || newSettings.contains("overlap_save");
        if (_plan != nullptr && !modeChanged && !overlap_save && newSettings.contains("b") && !b.empty()) {
            this->synchronizeStreams(); // queued chunks still read the old coefficients
            const bool referenceReallocates = b.size() > _refCapacity;
            if (gr4b200_fir_plan_set_taps(_plan, this->stream(), b.data(), b.size()) == GR4B200_OK) {
                if (referenceReallocates) {
                    _refCapacity = std::bit_ceil(b.size());
                    _stateOnly   = true;
                }
                return;
            }
        }
        if (modeChanged || newSettings.contains("b") || _plan == nullptr) {
            gr4b200_fir_plan_destroy(_plan);
            _plan        = nullptr; // re-created (history cleared) on the next chunk
            _refCapacity = detail::firReferenceCapacity(b.size());
            _stateOnly   = false;
        }
    }

    // the ring is asked for the past samples the reference's buffer holds, not just the nTaps - 1 the current `b` reads
    [[nodiscard]] std::size_t inputHistoryItems() const { return detail::firHistoryItems(std::max(b.size(), _refCapacity)); }
    [[nodiscard]] bool        historyInStream() { return !_stateOnly && _plan != nullptr && this->inputHistoryGranted() >= std::max(inputHistoryItems(), gr4b200_fir_plan_history_items(_plan)); }
    [[nodiscard]] bool        chunksIndependent() { return !_stateOnly && this->inputHistoryGranted() >= inputHistoryItems(); } // no carried state then

    void start() { // plan (taps + history in HBM) before the first chunk; a later `b` goes into the running plan (settingsChanged)
        if (_plan == nullptr && this->runsOnDevice()) {
            _plan = gr4b200_fir_plan_create(b.data(), b.size(), 1, planMode());
        }
    }

    gr::work::Status processBulk_cuda(void* stream, const T* input, T* output, std::size_t nIn, std::size_t /*nOut*/) {
        if (_plan == nullptr) {
            _plan = gr4b200_fir_plan_create(b.data(), b.size(), 1, planMode());
            if (_plan == nullptr) {
                return gr::work::Status::ERROR;
            }
        }
        // past samples: straight from the input ring when it keeps them (no state, no state kernel), else from the plan's state
        return detail::runFir(_plan, stream, input, output, nIn, historyInStream()) == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
    }

    gr4b200_fir_plan* _plan        = nullptr;
    std::size_t       _refCapacity = 32;    // capacity of the reference's HistoryBuffer for the coefficients seen so far
    bool              _stateOnly   = false; // a `b` that outgrew that buffer arrived mid-stream: the ring's past samples must not be read
};

// BasicFilterProto<T, Resampling<1,1,false>> with filter_type == FIR: designs its taps on every settings change and
// keeps every `decimate`-th output; input_chunk_size follows `decimate` (time_domain_filter.hpp:161-204).
template<typename T>
requires(std::is_same_v<T, float> || std::is_same_v<T, std::complex<float>>)
struct BasicDecimatingFilter : gr::Block<BasicDecimatingFilter<T>, gr::Resampling<1, 1, false>> {
    using gr::Block<BasicDecimatingFilter<T>, gr::Resampling<1, 1, false>>::Block;
    gr::PortIn<T>  in;
    gr::PortOut<T> out;
    FilterType     filter_type     = FilterType::FIR;
    Type           filter_response = Type::LOWPASS;
    gr::Size_t     filter_order    = 3;
    float          f_low           = 0.1f;
    float          f_high          = 0.2f;
    float          sample_rate     = 1.0f;
    gr::Size_t     decimate        = 1;
    gr::Size_t     fir_design_method = 11; // gr::algorithm::window::Type::Kaiser
    bool           exact           = true;
    GR_MAKE_REFLECTABLE(BasicDecimatingFilter, in, out, filter_type, filter_response, filter_order, f_low, f_high, sample_rate, decimate, fir_design_method, exact);

    ~BasicDecimatingFilter() { gr4b200_fir_plan_destroy(_plan); }

    void settingsChanged(const gr::property_map& /*oldSettings*/, const gr::property_map& /*newSettings*/) { designFilter(); }

    void designFilter() {
        if (filter_type != FilterType::FIR) {
            throw gr::exception("BasicDecimatingFilter: only FIR runs on the device (IIR feedback is sequential)");
        }
        this->input_chunk_size = decimate;
        std::vector<float> taps(1u << 16);
        const long         n = gr4b200_fir_design_f32_host(static_cast<int>(filter_response), filter_order, f_low, f_high, sample_rate, 1.0, 40.0, 1.6, static_cast<int>(fir_design_method), taps.data(), taps.size());
        if (n <= 0) {
            throw gr::exception("BasicDecimatingFilter: FIR design failed");
        }
        taps.resize(static_cast<std::size_t>(n));
        _taps = std::move(taps);
        if (_plan != nullptr) {
            // a running filter: the reference's new FilterImpl starts from a zeroed history (time_domain_filter.hpp:176-180), so
            // the past samples the input ring still holds must not be read any more -- the new plan's zeroed state is
            _stateOnly = true;
        }
        gr4b200_fir_plan_destroy(_plan);
        _plan = nullptr;
    }

    gr::work::Status processBulk_cuda(void* stream, const T* input, T* output, std::size_t nIn, std::size_t /*nOut*/) {
        if (_plan == nullptr) {
            _plan = gr4b200_fir_plan_create(_taps.data(), _taps.size(), decimate, exact ? GR4B200_FIR_EXACT : GR4B200_FIR_FAST);
            if (_plan == nullptr) {
                return gr::work::Status::ERROR;
            }
        }
        // past samples: straight from the input ring when it keeps them (no state, no state kernel), else from the plan's state
        const bool historyInStream = !_stateOnly && this->inputHistoryGranted() >= gr4b200_fir_plan_history_items(_plan);
        return detail::runFir(_plan, stream, input, output, nIn, historyInStream) == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
    }

    [[nodiscard]] std::size_t inputHistoryItems() const { return detail::firHistoryItems(_taps.size()); }
    [[nodiscard]] bool        chunksIndependent() { return !_stateOnly && this->inputHistoryGranted() >= inputHistoryItems(); }

    std::vector<float> _taps;
    gr4b200_fir_plan*  _plan      = nullptr;
    bool               _stateOnly = false; // re-designed mid-stream: history from the plan's (zeroed) state, not from the input ring
};

template<typename T>
requires std::is_same_v<T, std::complex<float>>
struct Decimator : gr::Block<Decimator<T>, gr::Resampling<1, 1, false>> {
    using gr::Block<Decimator<T>, gr::Resampling<1, 1, false>>::Block;
    gr::PortIn<T>  in;
    gr::PortOut<T> out;
    gr::Size_t     decim = 1;
    GR_MAKE_REFLECTABLE(Decimator, in, out, decim);
    void settingsChanged(const gr::property_map&, const gr::property_map&) { this->input_chunk_size = decim; }
    [[nodiscard]] bool chunksIndependent() const { return true; }
    gr::work::Status processBulk_cuda(void* stream, const T* input, T* output, std::size_t nIn, std::size_t /*nOut*/) {
        return gr4b200_decimate_cf32(stream, reinterpret_cast<const float*>(input), reinterpret_cast<float*>(output), nIn, decim) == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
    }
};

} // namespace gr::filter
