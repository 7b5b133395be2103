// TEST INFRASTRUCTURE ONLY: runs the kernels' per-thread arithmetic (gnuradio4_b200/csrc/{fir,fft}_core.cuh, the very
// functions the __global__ kernels call) on the HOST, thread by thread and tile by tile, so that the index mapping and
// the summation order can be checked against the oracle without a GPU (pytest -m "not gpu").
// Built by tests/test_host_emulation.py with: nvcc -std=c++17 -O1 -Xcompiler -fPIC,-ffp-contract=off -shared
#include <cstring>
#include <array>
#include <vector>
#include <atomic>
#include <thread>

#include "../gnuradio4_b200/csrc/fft_large.cuh"
#include "../gnuradio4_b200/csrc/fft_radix.cuh"
#include "../gnuradio4_b200/csrc/fir_core.cuh"
#include "../gnuradio4_b200/csrc/rotator_core.cuh"
#include "../gnuradio4_b200/csrc/sincos_core.cuh"

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
    static TapPairs pairs;
    if (nTaps <= kParamTaps) {
        fillTapPairs(pairs, taps, nTaps);
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
            if (sizeof(T) == 8 && nTaps <= kParamTaps) { // the kernels' parameter-pair route (fir_kernels.cuh useParamTaps)
                firTileThread<T, Threads, R, DLog2, Exact, Packed>(tid, tile.data(), layout, taps, pairs.pairs, nTaps, haloPad, tileStart, nOut, RoundingConsts{1.0f, -0.0f}, out);
            } else {
                firTileThread<T, Threads, R, DLog2, Exact>(tid, tile.data(), layout, taps, tapsT.data(), nTaps, haloPad, tileStart, nOut, RoundingConsts{1.0f, -0.0f}, out);
            }
        }
    }
}

template<typename T, bool Exact>
int dispatch(const float* taps, int nTaps, int decim, const T* in, T* out, long long nIn, const T* state) {
    switch (decim) { // same table as dispatchFir in fir.cu
    case 1:
        if (sizeof(T) == 8 && nIn <= kSmallCallTiles * 4096LL) {
            emulateFir<T, 256, kSmallCallR, 0, Exact>(taps, nTaps, in, out, nIn, state);
        } else {
            emulateFir<T, 256, kOutputsPerThreadD1<T>, 0, Exact>(taps, nTaps, in, out, nIn, state);
        }
        return 0;
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

// ---- FFT family (fft_radix.cuh): the passes exactly as fftRadixKernel sequences them, phase by phase ---------------------
namespace {
template<int N, int P>
void emulPasses(std::vector<Cx>& array, std::vector<Cx>& next, const float2* tables, Cx* dst) {
    using G = FftGeom<N>;
    for (int t = 0; t < G::kThreads; ++t) {
        Cx v[16];
        fftGather<N>(t, array.data(), v);
        fftPassCompute<N, P>(t, v, tables);
        if constexpr (P + 1 < G::kPasses) {
            fftScatter<N, P>(t, v, next.data());
        } else {
            for (int m = 0; m < 16; ++m) {
                dst[t + G::kThreads * m] = v[m];
            }
        }
    }
    if constexpr (P + 1 < G::kPasses) {
        array.swap(next);
        emulPasses<N, P + 1>(array, next, tables, dst);
    }
}

template<int N>
int emulFft(const float* in, float* out, long long batch, const float* window) {
    using G = FftGeom<N>;
    std::vector<float2> tables(G::kTableEntries > 0 ? G::kTableEntries : 1);
    fftFillTables<N>(tables.data());
    std::vector<float> windowT; // per-thread window layout exactly as gr4b200_fft_plan_create builds it
    if (window != nullptr) {
        windowT.resize(N);
        for (int t = 0; t < G::kThreads; ++t) {
            for (int m = 0; m < 16; ++m) {
                windowT[16 * t + m] = window[t + G::kThreads * m];
            }
        }
    }
    std::vector<Cx> array(G::kPadded), next(G::kPadded);
    for (long long xf = 0; xf < batch; ++xf) {
        const float2* src = reinterpret_cast<const float2*>(in) + xf * N;
        Cx*           dst = reinterpret_cast<Cx*>(out) + xf * N;
        for (int t = 0; t < G::kThreads; ++t) { // pass 0 reads the input (global or staged) at t + T m
            Cx v[16];
            for (int m = 0; m < 16; ++m) {
                v[m] = cxMake(src[t + G::kThreads * m].x, src[t + G::kThreads * m].y);
            }
            if (window != nullptr) {
                fftApplyWindow(t, windowT.data(), v);
            }
            fftPassCompute<N, 0>(t, v, tables.data());
            if constexpr (G::kPasses > 1) {
                fftScatter<N, 0>(t, v, array.data());
            } else {
                for (int m = 0; m < 16; ++m) {
                    dst[t + G::kThreads * m] = v[m];
                }
            }
        }
        if constexpr (G::kPasses > 1) {
            emulPasses<N, 1>(array, next, tables.data(), dst);
        }
    }
    return 0;
}

// worst bank-conflict degree (1 = conflict free) over all shared-memory access patterns of the kernel: 8-byte accesses
// are served per half-warp (16 lanes x 8 B = 128 B), 16-byte accesses per quarter-warp
template<int N>
int fftConflictDegree() {
    using G       = FftGeom<N>;
    constexpr int T = G::kThreads, lanesTotal = G::kCta;
    int           worst = 1;
    if (G::kPasses == 1) {
        return worst; // N = 16: one thread per transform, no shared memory
    }
    auto check8 = [&](auto&& addressOf) { // addressOf(threadIdx) -> element (8-byte) index in the CTA's array
        for (int base = 0; base < lanesTotal; base += 16) {
            int count[16] = {};
            for (int l = 0; l < 16; ++l) {
                const int c = ++count[addressOf(base + l) & 15];
                worst       = c > worst ? c : worst;
            }
        }
    };
    auto check16 = [&](auto&& addressOf) {
        for (int base = 0; base < lanesTotal; base += 8) {
            int count[8] = {};
            for (int l = 0; l < 8; ++l) {
                const int c = ++count[(addressOf(base + l) >> 1) & 7];
                worst       = c > worst ? c : worst;
            }
        }
    };
    for (int m = 0; m < 16; ++m) {
        check8([&](int tid) { const int i = tid % T + T * m; return (tid / T) * G::kPadded + i + (i >> 4); });       // gathers
        if constexpr (T >= 64) {
            check8([&](int tid) { return (tid / T) * N + tid % T + T * m; }); // staged input (bulk copy, N >= 1024 only)
        }
        check8([&](int tid) { const int i = fftScatterIndex<N, 0>(tid % T, m); return (tid / T) * G::kPadded + i + (i >> 4); });
        if constexpr (G::kPasses >= 3) {
            check8([&](int tid) { const int i = fftScatterIndex<N, 1>(tid % T, m); return (tid / T) * G::kPadded + i + (i >> 4); });
        }
        if constexpr (G::kPasses >= 4) {
            check8([&](int tid) { const int i = fftScatterIndex<N, 2>(tid % T, m); return (tid / T) * G::kPadded + i + (i >> 4); });
        }
        if constexpr (T >= 16) {
            check8([&](int tid) { return (tid / T) * G::kPadded + fftParkSlot(tid % T + T * m); }); // parking writes
        }
    }
    if constexpr (T >= 16) {
        // the hoisted forms used by the kernel (fftPark / fftParkReadBase) address exactly the slots of fftParkSlot
        std::vector<Cx> park(N, ~Cx{0});
        for (int t = 0; t < T; ++t) {
            Cx v[16];
            for (int m = 0; m < 16; ++m) {
                v[m] = static_cast<Cx>(t + T * m);
            }
            fftPark<N>(t, v, park.data());
        }
        for (int k = 0; k < N; ++k) {
            if (park[fftParkSlot(k)] != static_cast<Cx>(k)) {
                return -2;
            }
        }
        for (int t = 0; t < T; ++t) {
            for (int g = 0; g < 4; ++g) {
                for (int half = 0; half < 2; ++half) {
                    if (fftParkReadBase(t, half) + 4 * g * T != fftParkSlot(4 * (g * T + t) + 2 * half)) {
                        return -3;
                    }
                }
            }
        }
        for (int g = 0; g < 4; ++g) {
            for (int half = 0; half < 2; ++half) {
                check16([&](int tid) { return (tid / T) * G::kPadded + fftParkSlot(4 * (g * T + tid % T) + 2 * half); });
            }
        }
    }
    return worst;
}

// one pass of column transforms exactly as fftColumnKernel runs it: every phase for all threads of a tile, then the next
template<int L, bool First>
void emulColumns(const FftColumnArgs& a) {
    using G = FftColumnGeom<L>;
    std::vector<Cx>                smem(16 * G::kRegion);
    std::vector<std::array<Cx, 16>> regs(G::kThreads);
    const long long                tiles = a.batch * (a.cols / 16);
    for (long long tile = 0; tile < tiles; ++tile) {
        auto forAll = [&](auto&& phase) {
            for (int tid = 0; tid < G::kThreads; ++tid) {
                phase(tid, reinterpret_cast<Cx(&)[16]>(*regs[tid].data()));
            }
        };
        forAll([&](int tid, Cx(&v)[16]) { fftColumnPhaseLoad<L, First>(tid, tile, a, smem.data(), v); });
        forAll([&](int tid, Cx(&v)[16]) { fftColumnPhasePass<L, 1>(tid, a, smem.data(), v); });
        if constexpr (G::kPasses == 3) {
            forAll([&](int tid, Cx(&v)[16]) { fftColumnPhaseScatter<L, 1>(tid, v, smem.data()); });
            forAll([&](int tid, Cx(&v)[16]) { fftColumnPhasePass<L, 2>(tid, a, smem.data(), v); });
        }
        if constexpr (First) {
            forAll([&](int tid, Cx(&v)[16]) { fftColumnTwiddlePark<L>(tid, tile, a, v, smem.data()); });
            forAll([&](int tid, Cx(&)[16]) { fftColumnStoreTransposed<L>(tid, tile, a, smem.data()); });
        } else {
            forAll([&](int tid, Cx(&v)[16]) { fftColumnStoreRows<L>(tid, tile, a, v); });
        }
    }
}

template<bool First>
int emulColumnsOf(int length, const FftColumnArgs& a) {
    switch (length) {
    case 128: emulColumns<128, First>(a); return 0;
    case 256: emulColumns<256, First>(a); return 0;
    case 512: emulColumns<512, First>(a); return 0;
    default: return -1;
    }
}

template<int N>
void tablesOf(std::vector<float2>& tables) {
    tables.assign(FftGeom<N>::kTableEntries, make_float2(1.f, 0.f));
    fftFillTables<N>(tables.data());
}
void columnTables(int length, std::vector<float2>& tables) {
    switch (length) {
    case 128: return tablesOf<128>(tables);
    case 256: return tablesOf<256>(tables);
    default: return tablesOf<512>(tables);
    }
}

// worst bank-pair multiplicity (1 = conflict free) over the shared-memory access patterns of fftColumnKernel<L>: lanes
// run along the 16 columns of a tile, every column has its own region at an odd pitch; 8-byte accesses are served per
// half-warp. Also checks that a warp's transposed read touches exactly two 128-byte lines (returns 100 + lines if not).
template<int L>
int columnConflictDegree() {
    using G         = FftColumnGeom<L>;
    constexpr int T = G::kT, R = G::kRegion;
    int           worst = 1;
    auto check = [&](auto&& addressOf) {
        for (int base = 0; base < G::kThreads; base += 16) {
            int count[16] = {};
            for (int l = 0; l < 16; ++l) {
                const int c = ++count[addressOf(base + l) & 15];
                worst       = c > worst ? c : worst;
            }
        }
    };
    for (int m = 0; m < 16; ++m) {
        check([&](int tid) { return (tid & 15) * R + 17 * (tid >> 4) + m; });                                   // scatter after pass 0
        check([&](int tid) { const int t = tid >> 4; return (tid & 15) * R + (T >= 16 ? t + (t >> 4) + m * (T + T / 16) : t + T * m + ((T * m) >> 4)); }); // gather
        if (G::kPasses == 3) {
            check([&](int tid) { const int t = tid >> 4, b = (t / 16) * 256 + (t & 15); return (tid & 15) * R + b + (b >> 4) + m * 17; }); // scatter after pass 1
        }
        check([&](int tid) { return (tid & 15) * R + (tid >> 4) + T * m; });                                      // park (natural order)
        check([&](int tid) { return m * R + ((tid - m * (R % 16)) & (L - 1)); });                                 // transposed read of row m
        for (int warp = 0; warp < G::kThreads / 32; ++warp) {
            bool line[4096] = {};
            int  lines      = 0;
            for (int l = 0; l < 32; ++l) {
                const int a = (m * R + ((warp * 32 + l - m * (R % 16)) & (L - 1))) / 16;
                lines += line[a] ? 0 : 1;
                line[a] = true;
            }
            if (lines != 2 && !(warp == 0 || ((warp * 32 - m * (R % 16)) & (L - 1)) > L - 32)) { // the warp that wraps around the row may touch three
                return 100 + lines;
            }
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

#define GR4B200_FOR_EACH_FFT_SIZE(X) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048) X(4096) X(8192)

int emul_fft(int n, const float* in, float* out, long long batch, const float* window) {
    switch (n) {
#define X(N) \
    case N: return emulFft<N>(in, out, batch, window);
        GR4B200_FOR_EACH_FFT_SIZE(X)
#undef X
    default: return -1;
    }
}

// transforms of 8192 < n <= 262144 points: the two column passes of fft_large.cuh with the plan's tables
int emul_fft_large(int n, const float* in, float* out, long long batch, const float* window, int realInput) {
    if (n <= 8192 || n > kFftLargeMax || (n & (n - 1)) != 0) {
        return -1;
    }
    const int           n2 = fftLargeSecond(n), n1 = n / n2;
    std::vector<float2> tables1, tables2, twiddle(n);
    columnTables(n1, tables1);
    columnTables(n2, tables2);
    for (int j = 0; j < n; ++j) {
        const double angle = -2.0 * 3.14159265358979323846 * static_cast<double>(j) / static_cast<double>(n);
        twiddle[j]         = make_float2(static_cast<float>(std::cos(angle)), static_cast<float>(std::sin(angle)));
    }
    std::vector<Cx> scratch(static_cast<size_t>(n) * batch);
    FftColumnArgs   first{realInput != 0 ? nullptr : reinterpret_cast<const Cx*>(in), realInput != 0 ? in : nullptr, scratch.data(), window, twiddle.data(), tables1.data(), n2, batch, 0};
    if (emulColumnsOf<true>(n1, first) != 0) {
        return -1;
    }
    FftColumnArgs second{scratch.data(), nullptr, reinterpret_cast<Cx*>(out), nullptr, nullptr, tables2.data(), n1, batch, realInput};
    return emulColumnsOf<false>(n2, second);
}

int emul_fft_column_conflict_degree(int length) {
    switch (length) {
    case 128: return columnConflictDegree<128>();
    case 256: return columnConflictDegree<256>();
    case 512: return columnConflictDegree<512>();
    default: return -1;
    }
}

int emul_fft_conflict_degree(int n) {
    switch (n) {
#define X(N) \
    case N: return fftConflictDegree<N>();
        GR4B200_FOR_EACH_FFT_SIZE(X)
#undef X
    default: return -1;
    }
}

// the phase cycle rotator.cu builds once per plan (rotator_core.cuh findPhaseCycle) and the index rule its gather kernel
// applies: out[i] = phase in front of sample m[i]; returns 1 when the cycle closed (mu, lambda reported), 0 otherwise
int emul_rotator_cycle(float dphi, float startPhase, const unsigned long long* m, int count, float* out, unsigned long long* muOut, unsigned long long* lambdaOut) {
    std::vector<float> cycle;
    unsigned long long mu = 0, lambda = 0;
    if (!findPhaseCycle(dphi, startPhase, cycle, mu, lambda)) {
        return 0;
    }
    const unsigned long long size = mu + lambda;
    for (int i = 0; i < count; ++i) {
        const unsigned long long q = m[i] < size ? m[i] : mu + (m[i] - mu) % lambda;
        out[i]                     = cycle[q];
    }
    *muOut     = mu;
    *lambdaOut = lambda;
    return 1;
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

// sinCosGlibc (the mixer's cos/sin on the device) against the C library for every float whose bit pattern lies in
// [firstBits, lastBits), stepping by `stride`; returns the number of arguments where either result differs in any bit
// (NaN results only need to be NaN on both sides)
long long emul_sincos_mismatches(unsigned long long firstBits, unsigned long long lastBits, unsigned stride) {
    const unsigned                  nThreads = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<long long>          total{0};
    std::vector<std::thread>        pool;
    for (unsigned t = 0; t < nThreads; ++t) {
        pool.emplace_back([&, t] {
            long long bad = 0;
            for (unsigned long long b = firstBits + static_cast<unsigned long long>(t) * stride; b < lastBits; b += static_cast<unsigned long long>(nThreads) * stride) {
                const unsigned u = static_cast<unsigned>(b);
                float          y;
                std::memcpy(&y, &u, 4);
                volatile float yv = y;
                const float    sl = sinf(yv), cl = cosf(yv);
                float          s, c;
                gr4b200::sinCosGlibc(y, &s, &c);
                if (std::isnan(sl) || std::isnan(cl)) {
                    bad += !(std::isnan(s) && std::isnan(c));
                } else {
                    bad += std::memcmp(&s, &sl, 4) != 0 || std::memcmp(&c, &cl, 4) != 0;
                }
            }
            total += bad;
        });
    }
    for (auto& th : pool) {
        th.join();
    }
    return total.load();
}

} // extern "C"
