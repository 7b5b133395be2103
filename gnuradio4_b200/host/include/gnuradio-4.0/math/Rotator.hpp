// gr4b200 host layer -- gr::blocks::math::Rotator (reference: blocks/math/include/gnuradio-4.0/math/Rotator.hpp).
// Device only: the phase recurrence is replayed on the GPU by gr4b200_rotator_cf32 (see csrc/rotator.cu).
#pragma once

#include <complex>
#include <numbers>

#include "../Block.hpp"

namespace gr::blocks::math {

template<typename T>
requires std::is_same_v<T, std::complex<float>>
struct Rotator : gr::Block<Rotator<T>> {
    using gr::Block<Rotator<T>>::Block;
    using value_type = typename T::value_type;
    gr::PortIn<T>  in;
    gr::PortOut<T> out;
    gr::Annotated<float, "sample rate">           sample_rate     = 1.f;
    gr::Annotated<float, "frequency shift">       frequency_shift = 0.0f;
    gr::Annotated<value_type, "phase_increment">  phase_increment{0};
    gr::Annotated<value_type, "initial_phase">    initial_phase{0};
    GR_MAKE_REFLECTABLE(Rotator, in, out, sample_rate, frequency_shift, initial_phase, phase_increment);

    ~Rotator() { gr4b200_rotator_plan_destroy(_plan); }

    void settingsChanged(const gr::property_map& /*oldSettings*/, const gr::property_map& newSettings) {
        const bool hasShift = newSettings.contains("frequency_shift"), hasIncrement = newSettings.contains("phase_increment");
        if (hasShift && hasIncrement) {
            throw gr::exception("cannot set both 'frequency_shift' and 'phase_increment' in new setting (XOR)");
        }
        if (hasShift) {
            phase_increment = gr4b200_rotator_phase_increment(frequency_shift, sample_rate);
        } else if (hasIncrement) {
            frequency_shift = static_cast<float>(phase_increment.value / (value_type(2) * std::numbers::pi_v<value_type>)) * sample_rate;
        }
        _dirty = true; // the accumulated phase restarts at initial_phase (Rotator.hpp:48)
    }

    gr::work::Status processBulk_cuda(void* stream, const T* input, T* output, std::size_t nIn, std::size_t /*nOut*/) {
        if (_dirty || _plan == nullptr) {
            gr4b200_rotator_plan_destroy(_plan);
            _plan  = gr4b200_rotator_plan_create(phase_increment, initial_phase);
            _dirty = false;
            if (_plan == nullptr) {
                return gr::work::Status::ERROR;
            }
        }
        return gr4b200_rotator_cf32(_plan, stream, reinterpret_cast<const float*>(input), reinterpret_cast<float*>(output), nIn) == GR4B200_OK ? gr::work::Status::OK : gr::work::Status::ERROR;
    }

    [[nodiscard]] float accumulatedPhase() const { return _plan != nullptr ? gr4b200_rotator_get_phase(_plan) : static_cast<float>(initial_phase.value); }

    gr4b200_rotator_plan* _plan  = nullptr;
    bool                  _dirty = true;
};

} // namespace gr::blocks::math
