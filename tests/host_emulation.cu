// TEST INFRASTRUCTURE ONLY: runs the kernels' per-thread arithmetic (gnuradio4_b200/csrc/{fir,fft}_core.cuh, the very
// functions the __global__ kernels call) on the HOST, thread by thread and tile by tile, so that the index mapping and
// the summation order can be checked against the oracle without a GPU (pytest -m "not gpu").
// Built by tests/test_host_emulation.py with: nvcc -std=c++17 -O1 -Xcompiler -fPIC,-ffp-contract=off -shared
#include <cstring>
#include <vector>

#include "../gnuradio4_b200/csrc/fft_core.cuh"
#include "../gnuradio4_b200/csrc/fir_core.cuh"
#include "../gnuradio4_b200/csrc/rotator_core.cuh"

using namespace gr4b200;

namespace {
template<typename T, int Threads, int R, int DLog2, bool Exact>
void emulateFir(const float* taps, int nTaps, const T* in, T* out, long long nIn, const T* state) {
    using Cfg               = FirConfig<T, Threads, R, DLog2, Exact>;
    using Layout            = TileLayout<T, DLog2>;
    const int       haloPad = (nTaps - 1 + 15) / 16 * 16;
    const long long nOut    = nIn >> DLog2;
    const long long nTiles  = (nIn + Cfg::TileIn - 1) / Cfg::TileIn;
    const int       extended = haloPad + Cfg::TileIn;
    const Layout    layout{Layout::pitchFor(extended)};
    std::vector<T>  tile(static_cast<size_t>(Cfg::D) * layout.pitch); // staged exactly as the kernels stage it (phase major for D > 1)
    const int          pitch = lanePitchFor(nTaps);
    std::vector<float> tapsT(static_cast<size_t>(kLanes) * pitch + 8, 0.f); // lane-major copy exactly as the kernel prologue builds it
    for (int k = 0; k < kLanes * pitch; ++k) {
        const int j = k / pitch, m = k % pitch;
        tapsT[k]    = j + kLanes * m < nTaps ? taps[j + kLanes * m] : 0.f;
    }
    for (long long t = 0; t < nTiles; ++t) {
        const long long tileStart = t * Cfg::TileIn;
        for (int e = 0; e < extended; ++e) {
            const long long q = tileStart - haloPad + e;
            T               v = zeroOf(T{});
            if (q < 0) {
                v = state != nullptr ? state[haloPad + q] : zeroOf(T{});
            } else if (q < nIn) {
                v = in[q];
            }
            tile[layout(e)] = v;
        }
        for (int tid = 0; tid < Threads; ++tid) {
            firTileThread<T, Threads, R, DLog2, Exact>(tid, tile.data(), layout, taps, tapsT.data(), nTaps, haloPad, tileStart, nOut, RoundingConsts{1.0f, -0.0f}, out);
        }
    }
}

template<typename T, bool Exact>
int dispatch(const float* taps, int nTaps, int decim, const T* in, T* out, long long nIn, const T* state) {
    switch (decim) { // same table as dispatchFir in fir.cu
    case 1: emulateFir<T, 256, kOutputsPerThreadD1<T>, 0, Exact>(taps, nTaps, in, out, nIn, state); return 0;
    case 2: emulateFir<T, kDecimThreads2, kDecimR2, 1, Exact>(taps, nTaps, in, out, nIn, state); return 0;
    case 4: emulateFir<T, kDecimThreads4, kDecimR4, 2, Exact>(taps, nTaps, in, out, nIn, state); return 0;
    case 8: emulateFir<T, kDecimThreads8, kDecimR8, 3, Exact>(taps, nTaps, in, out, nIn, state); return 0;
    case 16: emulateFir<T, kDecimThreads16, kDecimR16, 4, Exact>(taps, nTaps, in, out, nIn, state); return 0;
    default: return -1;
    }
}

// shared-memory wavefronts of the window loads of one warp: for every lane j and window entry, the (half-)warp's
// element addresses are mapped to banks; returns the worst number of wavefronts beyond the minimum (0 = conflict free)
template<typename T, int Threads, int R, int DLog2>
int firBankConflicts(int nTaps) {
    using Cfg         = FirConfig<T, Threads, R, DLog2, true>;
    using Layout      = TileLayout<T, DLog2>;
    const int haloPad = (nTaps - 1 + 15) / 16 * 16;
    const Layout layout{Layout::pitchFor(haloPad + Cfg::TileIn)};
    constexpr int wordsPerElem = sizeof(T) / 4;
    constexpr int group        = 32 / wordsPerElem; // threads served by one wavefront when conflict free
    constexpr int Step         = kLanes >> DLog2;
    int worst = 0;
    for (int base = 0; base + group <= Threads; base += group) {
        for (int j = 0; j < kLanes; ++j) {
            for (int q = -7; q < R; ++q) {
                int perBank[32] = {0};
                for (int tid = base; tid < base + group; ++tid) {
                    const int seg = tid / Cfg::G, tsub = tid % Cfg::G;
                    const int e0  = haloPad + seg * (kLanes * R) + tsub * Cfg::D;
                    const int idx = layout(e0 - j) + Step * q;
                    for (int w = 0; w < wordsPerElem; ++w) {
                        ++perBank[(idx * wordsPerElem + w) % 32];
                    }
                }
                for (int b = 0; b < 32; ++b) {
                    worst = perBank[b] - 1 > worst ? perBank[b] - 1 : worst;
                }
            }
        }
        // staging writes: consecutive threads write consecutive extended samples
        int perBank[32] = {0};
        for (int tid = 0; tid < group; ++tid) {
            const int idx = layout(base + tid);
            for (int w = 0; w < wordsPerElem; ++w) {
                ++perBank[(idx * wordsPerElem + w) % 32];
            }
        }
        for (int b = 0; b < 32; ++b) {
            worst = perBank[b] - 1 > worst ? perBank[b] - 1 : worst;
        }
    }
    return worst;
}
} // namespace

extern "C" {

// state: haloPad samples preceding in[0] (may be NULL = zeros); complex = 1 -> float2 stream
int emul_fir(const float* taps, int nTaps, int decim, int exact, int complexStream, const float* in, float* out, long long nIn, const float* state) {
    if (complexStream) {
        auto* i = reinterpret_cast<const float2*>(in);
        auto* o = reinterpret_cast<float2*>(out);
        auto* s = reinterpret_cast<const float2*>(state);
        return exact ? dispatch<float2, true>(taps, nTaps, decim, i, o, nIn, s) : dispatch<float2, false>(taps, nTaps, decim, i, o, nIn, s);
    }
    return exact ? dispatch<float, true>(taps, nTaps, decim, in, out, nIn, state) : dispatch<float, false>(taps, nTaps, decim, in, out, nIn, state);
}

// worst extra shared-memory wavefronts of the FIR tile accesses for the dispatch table's configuration (0 = conflict free)
int emul_fir_bank_conflicts(int nTaps, int decim, int complexStream) {
    switch (decim) {
    case 1: return complexStream ? firBankConflicts<float2, 256, kOutputsPerThreadD1<float2>, 0>(nTaps) : firBankConflicts<float, 256, kOutputsPerThreadD1<float>, 0>(nTaps);
    case 2: return complexStream ? firBankConflicts<float2, kDecimThreads2, kDecimR2, 1>(nTaps) : firBankConflicts<float, kDecimThreads2, kDecimR2, 1>(nTaps);
    case 4: return complexStream ? firBankConflicts<float2, kDecimThreads4, kDecimR4, 2>(nTaps) : firBankConflicts<float, kDecimThreads4, kDecimR4, 2>(nTaps);
    case 8: return complexStream ? firBankConflicts<float2, kDecimThreads8, kDecimR8, 3>(nTaps) : firBankConflicts<float, kDecimThreads8, kDecimR8, 3>(nTaps);
    case 16: return complexStream ? firBankConflicts<float2, kDecimThreads16, kDecimR16, 4>(nTaps) : firBankConflicts<float, kDecimThreads16, kDecimR16, 4>(nTaps);
    default: return -1;
    }
}

int emul_fft4096(const float* in, float* out, long long batch, const float* window) {
    std::vector<float2> powers1(4 * 256), powers2(4 * 16), sA(kN4096), sB(256 * kRowStride4096);
    fillPowerTable(powers1.data(), 256, 4096);
    fillPowerTable(powers2.data(), 16, 256);
    std::vector<float> windowT; // per-thread window layout exactly as gr4b200_fft_plan_create builds it
    if (window != nullptr) {
        windowT.resize(4096);
        for (int t = 0; t < 256; ++t) {
            for (int n1 = 0; n1 < 16; ++n1) {
                windowT[16 * t + n1] = window[256 * n1 + t];
            }
        }
    }
    for (long long xf = 0; xf < batch; ++xf) {
        const float2* src = reinterpret_cast<const float2*>(in) + xf * kN4096;
        float2*       dst = reinterpret_cast<float2*>(out) + xf * kN4096;
        for (int t = 0; t < 256; ++t) {
            float2 x[16];
            fft4096Pass1(t, src, window != nullptr ? windowT.data() : nullptr, powers1.data(), x);
            fft4096Store1(t, x, sA.data());
        }
        for (int t = 0; t < 256; ++t) {
            fft4096Pass2(t, sA.data(), powers2.data(), sB.data());
        }
        for (int t = 0; t < 256; ++t) {
            float2 x[16];
            fft4096Pass3(t, sB.data(), x);
            for (int k3 = 0; k3 < 16; ++k3) {
                dst[k3 * 256 + t] = x[k3];
            }
        }
    }
    return 0;
}

int emul_fft256(const float* in, float* out, long long batch, const float* window) {
    std::vector<float2> powers1(4 * 16), sRow(16 * 17);
    fillPowerTable(powers1.data(), 16, 256);
    for (long long xf = 0; xf < batch; ++xf) {
        const float2* src = reinterpret_cast<const float2*>(in) + xf * kN256;
        float2*       dst = reinterpret_cast<float2*>(out) + xf * kN256;
        for (int t = 0; t < 16; ++t) {
            float2 x[16];
            fft256Pass1(t, src, window, powers1.data(), x);
            fft256Store1(t, x, sRow.data());
        }
        for (int t = 0; t < 16; ++t) {
            float2 x[16];
            fft256Pass2(t, sRow.data(), x);
            for (int k2 = 0; k2 < 16; ++k2) {
                dst[k2 * 16 + t] = x[k2];
            }
        }
    }
    return 0;
}

// mixer phase lookup exactly as rotator.cu performs it (prefix, base table, lifting, lookup + residual replay):
// out[i] = phase in front of sample index m[i] for a call of nSamples samples starting at startPhase.
// returns the number of landing states, 0 if this dphi takes the serial path, -1 if the grid assumption was violated
int emul_rotator_phases(float dphi, float startPhase, unsigned long long nSamples, const unsigned long long* m, int count, float* out) {
    Landing l{};
    if (!landingFor(dphi, l)) {
        return 0;
    }
    const int                       levels = liftingLevels(dphi, nSamples);
    std::vector<unsigned long long> tables(static_cast<size_t>(levels) * l.nStates);
    bool                            violated = false;
    for (int k = 0; k < l.nStates; ++k) {
        tables[k] = baseTableEntry(l, k, 1ull << 26, violated);
    }
    if (violated) {
        return -1;
    }
    for (int j = 1; j < levels; ++j) {
        for (int k = 0; k < l.nStates; ++k) {
            tables[static_cast<size_t>(j) * l.nStates + k] = liftTableEntry(tables.data() + static_cast<size_t>(j - 1) * l.nStates, k);
        }
    }
    Prefix prefix{};
    computePrefix(l, startPhase, nSamples, &prefix);
    for (int i = 0; i < count; ++i) {
        out[i] = phaseBeforeSample(l, startPhase, prefix, tables.data(), levels, m[i]);
    }
    return l.nStates;
}

} // extern "C"
