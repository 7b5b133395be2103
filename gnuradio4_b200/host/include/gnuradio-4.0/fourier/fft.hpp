// gr4b200 host layer -- gr::blocks::fft::FFT for std::complex<float> and float input
// (reference: blocks/fourier/include/gnuradio-4.0/fourier/fft.hpp:29-250). The reference emits one DataSet per chunk of
// fftSize samples; here one output item is the DataSet's signal_values block: 4 x M floats, M = fftSize for complex
// input {Magnitude (fft-shifted), Phase (fft-shifted), Re, Im} and M = fftSize / 2 for real input {Magnitude and Phase of
// bins [0, M), Re and Im of the last M bins of the spectrum, fft.hpp:212-217}, resident in HBM. `materialise()` builds the host-side
// DataSet-shaped view (axis, names, units, ranges) only when a host consumer asks for it.
#pragma once

#include <array>
#include <complex>
#include <string>
#include <vector>

#include "../Block.hpp"

namespace gr::blocks::fft {

template<std::size_t N>
struct SpectrumFrame { // signal_values of one DataSet
    std::array<float, N> magnitude, phase, re, im;
};

struct DataSetView { // the parts of gr::DataSet<float> that FFT::createDataset fills (fft.hpp:173-250)
    std::vector<std::string>          signal_names, signal_quantities, signal_units;
    std::vector<float>                axis_values;
    std::vector<float>                signal_values; // 4 * N
    std::vector<std::array<float, 2>> signal_ranges;
};

template<typename T, std::size_t FftSize = 4096>
requires(std::is_same_v<T, std::complex<float>> || std::is_same_v<T, float>)
struct FFT : gr::Block<FFT<T, FftSize>, gr::Resampling<FftSize, 1>> {
    using gr::Block<FFT<T, FftSize>, gr::Resampling<FftSize, 1>>::Block;
    static constexpr bool        computeFullSpectrum = std::is_same_v<T, std::complex<float>>; // fft.hpp:123
    static constexpr std::size_t kBins               = computeFullSpectrum ? FftSize : FftSize / 2;
    using Frame = SpectrumFrame<kBins>;
    gr::PortIn<T>      in;
    gr::PortOut<Frame> out;
    gr::Size_t         fftSize     = FftSize; // fixed at compile time in this layer (the output item is sized by it)
    std::string        window      = "Hann";
    bool               outputInDb  = false;
    bool               outputInDeg = false;
    bool               unwrapPhase = false;
    float              sample_rate = 1.f;
    std::string        signal_name = "unknown signal";
    std::string        signal_unit = "a.u.";
    GR_MAKE_REFLECTABLE(FFT, in, out, fftSize, window, outputInDb, outputInDeg, unwrapPhase, sample_rate, signal_name, signal_unit);

    ~FFT() { gr4b200_fft_plan_destroy(_plan); }

    void settingsChanged(const gr::property_map& /*oldSettings*/, const gr::property_map& newSettings) {
        if (fftSize != FftSize) {
            throw gr::exception("FFT: fftSize is a template parameter of this block (FFT<T, N>)");
        }
        if (newSettings.contains("window") || _plan == nullptr) {
            static constexpr std::array<std::string_view, 12> names{"none", "rectangular", "hamming", "hann", "hannexp", "blackman", "nuttall", "blackmanharris", "blackmannuttall", "flattop", "exponential", "kaiser"};
            std::string lower = window;
            for (auto& c : lower) {
                c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
            }
            for (std::size_t i = 0; i < names.size(); ++i) {
                if (names[i] == lower) {
                    _windowType = static_cast<int>(i); // unknown names keep the previous type (fft.hpp:135)
                }
            }
            gr4b200_fft_plan_destroy(_plan);
            _plan = nullptr;
        }
    }

    [[nodiscard]] bool chunksIndependent() const { return true; } // every transform stands alone

    bool createPlan() {
        std::vector<float> w(FftSize);
        if (gr4b200_window_f32_host(_windowType, FftSize, 1.6f, w.data()) != GR4B200_OK) {
            return false;
        }
        _plan = gr4b200_fft_plan_create(FftSize, w.data());
        return _plan != nullptr;
    }

    void start() { // twiddle tables and window in HBM before the first chunk
        if (_plan == nullptr && this->runsOnDevice()) {
            createPlan();
        }
    }

    gr::work::Status processBulk_cuda(void* stream, const T* input, Frame* output, std::size_t nIn, std::size_t nOut) {
        if (_plan == nullptr && !createPlan()) {
            return gr::work::Status::ERROR;
        }
        const unsigned flags = (outputInDb ? GR4B200_FFT_OUTPUT_IN_DB : 0u) | (outputInDeg ? GR4B200_FFT_OUTPUT_IN_DEG : 0u) | (unwrapPhase ? GR4B200_FFT_UNWRAP_PHASE : 0u);
        (void)nIn;
        int rc;
        if constexpr (computeFullSpectrum) {
            rc = gr4b200_fft_block_cf32(_plan, stream, reinterpret_cast<const float*>(input), nOut, flags, reinterpret_cast<float*>(output), nullptr);
        } else {
            rc = gr4b200_fft_block_f32(_plan, stream, input, nOut, flags, reinterpret_cast<float*>(output), nullptr);
        }
        return rc == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
    }

    // host-side DataSet view of one frame that has been copied back (createDataset, fft.hpp:173-250)
    [[nodiscard]] DataSetView materialise(const Frame& frame) const {
        DataSetView ds;
        ds.signal_names      = {"Magnitude(" + signal_name + ")", "Phase(" + signal_name + ")", "Re(FFT(" + signal_name + "))", "Im(FFT(" + signal_name + "))"};
        ds.signal_quantities = {"Magnitude(FFT)", "Phase(FFT)", "Re(FFT)", "Im(FFT)"};
        ds.signal_units      = {signal_unit + "/√Hz", "rad", "Re" + signal_unit, "Im" + signal_unit};
        const float width    = sample_rate / static_cast<float>(FftSize);
        const float offset   = computeFullSpectrum ? static_cast<float>(kBins / 2) * width : 0.f; // real input: [DC, +fs/2) (fft.hpp:193-196)
        ds.axis_values.resize(kBins);
        for (std::size_t i = 0; i < kBins; ++i) {
            ds.axis_values[i] = static_cast<float>(i) * width - offset;
        }
        ds.signal_values.reserve(4 * kBins);
        for (const auto* plane : {&frame.magnitude, &frame.phase, &frame.re, &frame.im}) {
            ds.signal_values.insert(ds.signal_values.end(), plane->begin(), plane->end());
            const auto [lo, hi] = std::minmax_element(plane->begin(), plane->end());
            ds.signal_ranges.push_back({*lo, *hi});
        }
        return ds;
    }

    gr4b200_fft_plan* _plan       = nullptr;
    int               _windowType = 3; // Hann
};

} // namespace gr::blocks::fft
