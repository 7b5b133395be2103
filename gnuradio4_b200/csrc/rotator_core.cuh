// Mixer phase machinery shared by the device kernels (rotator.cu) and the host emulation used by the CPU-side tests
// (tests/host_emulation.cu). See rotator.cu for the method.
#pragma once

#include <cuda_runtime.h>

#include <cmath>

#ifndef GR4B200_HD
#define GR4B200_HD __host__ __device__ __forceinline__
#endif

namespace gr4b200 {

constexpr float kTwoPi      = 6.283185307179586476925286766559f; // 2.f * pi_v<float>
constexpr int   kTile       = 4096;                               // samples per tile
constexpr int   kRun        = 8;                                  // samples per checkpoint (one column of the /8 phase-major tile: the fused DDC replays nothing it does not use)
constexpr int   kRunsPerTile = kTile / kRun;
constexpr unsigned long long kStepsSaturated = (1ull << 39);      // "more steps than any call will ask for"
constexpr int   kNextBits   = 24;
constexpr unsigned long long kNextMask = (1ull << kNextBits) - 1;

GR4B200_HD float stepPhase(float phase, float dphi, bool& wrapped) {
#ifdef __CUDA_ARCH__
    // branch free: threads of a warp wrap at different samples, a real branch would diverge on almost every step
    phase              = __fadd_rn(phase, dphi);
    const float down   = __fsub_rn(phase, kTwoPi);
    const float up     = __fadd_rn(phase, kTwoPi);
    const bool  isOver = phase > kTwoPi;
    const bool  isNeg  = phase < 0.f;
    wrapped            = isOver || isNeg;
    phase              = isOver ? down : (isNeg ? up : phase);
#else
    phase += dphi;
    wrapped = false;
    if (phase > kTwoPi) {
        phase -= kTwoPi;
        wrapped = true;
    } else if (phase < 0.f) {
        phase += kTwoPi;
        wrapped = true;
    }
#endif
    return phase;
}

// The same step when the phase is known to lie in [0, 2 pi_f] in front of it (every phase after the first wrap does):
// with dphi > 0 the sum cannot be negative, with dphi < 0 (|dphi| <= pi) it cannot exceed 2 pi_f, so one of the two
// tests -- and the value it would select -- drops out. Same bits as stepPhase for such phases.
template<bool Positive>
GR4B200_HD float stepPhaseInRange(float phase, float dphi) {
#ifdef __CUDA_ARCH__
    phase = __fadd_rn(phase, dphi);
    if constexpr (Positive) {
        const float down = __fsub_rn(phase, kTwoPi);
        return phase > kTwoPi ? down : phase;
    } else {
        const float up = __fadd_rn(phase, kTwoPi);
        return phase < 0.f ? up : phase;
    }
#else
    bool wrapped;
    return stepPhase(phase, dphi, wrapped);
#endif
}

struct Landing { // description of the landing-state grid for one dphi
    float dphi;
    float grid;      // 2^-21 (dphi > 0) or 2^-22 (dphi < 0)
    int   nStates;   // K
    int   positive;  // dphi > 0
};

GR4B200_HD float stateToPhase(const Landing& l, int k) { return l.positive ? k * l.grid : kTwoPi - k * l.grid; }
// index of a just-wrapped phase, or -1 if it is not a landing state (only possible before the orbit has settled)
GR4B200_HD int phaseToState(const Landing& l, float phase) {
    const float offset = l.positive ? phase : kTwoPi - phase;
    const float scaled = offset / l.grid; // exact: power-of-two grid
    if (!(scaled >= 0.f) || scaled >= static_cast<float>(l.nStates) || scaled != floorf(scaled)) {
        return -1;
    }
    return static_cast<int>(scaled);
}

struct Prefix {               // written by prefixKernel
    unsigned long long steps;  // samples consumed until the first landing (or the whole call if it never lands)
    int                state;  // landing state index, -1 if none was reached within the call
    float              phase;  // phase after `steps` samples
};

// replay from the call's start phase until the first landing state
GR4B200_HD void computePrefix(const Landing& l, float startPhase, unsigned long long nSamples, Prefix* prefix) {
    float              phase = startPhase;
    unsigned long long steps = 0;
    int                state = -1;
    while (steps < nSamples) {
        bool wrapped;
        phase = stepPhase(phase, l.dphi, wrapped);
        ++steps;
        if (wrapped) {
            state = phaseToState(l, phase);
            if (state >= 0) {
                break;
            }
        }
    }
    prefix->steps = steps;
    prefix->state = state;
    prefix->phase = phase;
}

// T0[k]: replay one revolution from landing state k. entry = steps << 24 | next. A stalled accumulator or a revolution
// longer than maxSteps saturates the step count (such an entry is never taken by the lookup).
GR4B200_HD unsigned long long baseTableEntry(const Landing& l, int k, unsigned long long maxSteps, bool& gridViolated) {
    float              phase = stateToPhase(l, k);
    unsigned long long steps = 0;
    int                next  = -1;
    while (steps < maxSteps) {
        bool        wrapped;
        const float before = phase;
        phase              = stepPhase(phase, l.dphi, wrapped);
        ++steps;
        if (wrapped) {
            next = phaseToState(l, phase);
            if (next < 0) {
                gridViolated = true; // grid assumption violated: host falls back to the serial replay
                next         = 0;
            }
            break;
        }
        if (phase == before) { // accumulator stalled: it will never wrap
            break;
        }
    }
    if (next < 0) {
        return (kStepsSaturated << kNextBits) | static_cast<unsigned long long>(k);
    }
    return (steps << kNextBits) | static_cast<unsigned long long>(next);
}

GR4B200_HD unsigned long long liftTableEntry(const unsigned long long* prev, int k) {
    const unsigned long long a     = prev[k];
    const unsigned long long b     = prev[a & kNextMask];
    unsigned long long       steps = (a >> kNextBits) + (b >> kNextBits);
    steps                          = steps > kStepsSaturated ? kStepsSaturated : steps;
    return (steps << kNextBits) | (b & kNextMask);
}

// phase in front of sample index m (i.e. after m steps from the call's start phase)
GR4B200_HD float phaseBeforeSample(const Landing& l, float startPhase, const Prefix& prefix, const unsigned long long* tables, int nLevels, unsigned long long m) {
    if (m <= prefix.steps || prefix.state < 0) {
        if (m == prefix.steps) {
            return prefix.phase;
        }
        float phase = startPhase; // inside the shared prefix (short): plain replay
        for (unsigned long long i = 0; i < m; ++i) {
            bool wrapped;
            phase = stepPhase(phase, l.dphi, wrapped);
        }
        return phase;
    }
    unsigned long long remaining = m - prefix.steps;
    int                k         = prefix.state;
    for (int j = nLevels - 1; j >= 0; --j) {
        const unsigned long long e     = tables[static_cast<size_t>(j) * l.nStates + k];
        const unsigned long long steps = e >> kNextBits;
        if (steps <= remaining) {
            remaining -= steps;
            k = static_cast<int>(e & kNextMask);
        }
    }
    // the level-0 entry may still fit several times (levels are capped): walk single revolutions, then replay the rest
    while (true) {
        const unsigned long long e     = tables[k];
        const unsigned long long steps = e >> kNextBits;
        if (steps > remaining) {
            break;
        }
        remaining -= steps;
        k = static_cast<int>(e & kNextMask);
    }
    float phase = stateToPhase(l, k);
    for (unsigned long long i = 0; i < remaining; ++i) {
        bool wrapped;
        phase = stepPhase(phase, l.dphi, wrapped);
    }
    return phase;
}

// landing-state grid for dphi; false => |dphi| > pi, zero or non-finite: callers use the serial replay
inline bool landingFor(float dphi, Landing& l) {
    if (!std::isfinite(dphi) || dphi == 0.f || std::fabs(dphi) > 3.14159265358979f) {
        return false;
    }
    l.dphi     = dphi;
    l.positive = dphi > 0.f ? 1 : 0;
    l.grid     = l.positive ? 1.f / 2097152.f : 1.f / 4194304.f; // 2^-21 : 2^-22
    l.nStates  = static_cast<int>(std::floor(std::fabs(static_cast<double>(dphi)) / l.grid)) + 2;
    return l.nStates < (1 << kNextBits);
}

// number of lifting levels so that 2^(levels-1) revolutions cover nSamples
inline int liftingLevels(float dphi, unsigned long long nSamples) {
    double minSteps = std::floor(6.283185307179586 / std::fabs(static_cast<double>(dphi))) - 1.0;
    minSteps        = minSteps < 1.0 ? 1.0 : minSteps;
    int levels      = 1;
    while (std::ldexp(minSteps, levels - 1) < static_cast<double>(nSamples) && levels < 40) {
        ++levels;
    }
    return levels;
}

} // namespace gr4b200
