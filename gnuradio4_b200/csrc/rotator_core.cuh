// Mixer phase machinery shared by the device kernels (rotator.cu) and the host emulation used by the CPU-side tests
// (tests/host_emulation.cu). See rotator.cu for the method.
#pragma once

#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <vector>

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


// The phase recurrence is a map on the 2^32 float patterns, so a plan's phase sequence is eventually periodic. Replays it on
// the host (same float operations as the device: add, compare, add) from `startPhase` until it closes: cycle[i] = phase in
// front of sample i for i < mu + lambda, and the sequence repeats with period lambda from sample mu on. Only the phases right
// after a wrap are remembered (one per revolution); a stalled accumulator closes with lambda = 1. false: not closed within
// 40 M samples (cannot happen for |dphi| <= pi, where landing states x steps per revolution <= 2 pi * 2^22 = 26 M).
inline bool findPhaseCycle(float dphi, float startPhase, std::vector<float>& cycle, unsigned long long& mu, unsigned long long& lambda) {
    constexpr unsigned long long kMaxSamples = 40ull << 20;
    // open addressing over the phases seen right after a wrap: at most |dphi| * 2^22 + 2 landing states for |dphi| <= pi,
    // so the table is sized from the increment (2^10 .. 2^24 slots, at most half full)
    const double       bound  = std::fabs(static_cast<double>(dphi)) <= 3.1415927 ? std::fabs(static_cast<double>(dphi)) * 4194304.0 + 1024.0 : 8388608.0;
    int                log2Slots = 10;
    while ((1ull << log2Slots) < 2.0 * bound && log2Slots < 24) {
        ++log2Slots;
    }
    const unsigned long long        slots = 1ull << log2Slots;
    std::vector<unsigned>           keys(slots, 0u);      // phase bits + 1 (0 = empty; +1 cannot overflow: 0xffffffff is a NaN)
    std::vector<unsigned long long> seenAt(slots, 0ull);  // sample index in front of which that phase stood
    cycle.clear();
    cycle.reserve(1u << 22);
    float              phase = startPhase;
    unsigned long long wraps = 0;
    mu = lambda = 0;
    for (unsigned long long i = 0; i < kMaxSamples && lambda == 0; ++i) {
        cycle.push_back(phase);
        bool        wrapped = false;
        const float next    = stepPhase(phase, dphi, wrapped);
        unsigned    nextBits, phaseBits;
        std::memcpy(&nextBits, &next, sizeof nextBits);
        std::memcpy(&phaseBits, &phase, sizeof phaseBits);
        if (nextBits == phaseBits) { // the accumulator no longer moves: period 1 from here
            mu = i, lambda = 1;
        } else if (wrapped) {
            if (++wraps > slots / 2) {
                return false;
            }
            const unsigned     key  = nextBits + 1u;
            unsigned long long slot = (static_cast<unsigned long long>(nextBits) * 0x9E3779B97F4A7C15ull) >> (64 - log2Slots);
            while (keys[slot] != 0u && keys[slot] != key) {
                slot = (slot + 1) & (slots - 1);
            }
            if (keys[slot] == key) { // this phase stood in front of sample seenAt[slot] before: the sequence repeats from there
                mu = seenAt[slot], lambda = i + 1 - seenAt[slot];
            } else {
                keys[slot] = key, seenAt[slot] = i + 1;
            }
        }
        phase = next;
    }
    return lambda != 0 && cycle.size() == mu + lambda;
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
