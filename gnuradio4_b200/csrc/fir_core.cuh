// Per-thread FIR arithmetic shared by the device kernel (fir.cu) and the host emulation used by the CPU-side tests
// (tests/host_emulation.cu). See fir.cu for the contract and the thread mapping.
#pragma once

#include <cuda_runtime.h>

#ifndef GR4B200_HD
#define GR4B200_HD __host__ __device__ __forceinline__
#endif

namespace gr4b200 {

constexpr int kLanes = 16; // 64-byte PSTL lane block / sizeof(float)

// round-to-nearest single operations that the compiler may not contract into FMAs
GR4B200_HD float fmulRn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b; // host emulation is built with -ffp-contract=off
#endif
}
GR4B200_HD float faddRn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
GR4B200_HD float fmaRn(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

// ---- vector element helpers -----------------------------------------------------------------------------------------
// A complex<float> sample is handled as one packed f32x2 register pair: Blackwell's FMUL2 / FADD2 issue at the same rate
// as scalar FMUL / FADD but produce two results, so "round the product, then round the sum" (the reference's arithmetic)
// costs the same pipe time as a fused multiply-add (measured: 36.9 T mul+add pairs/s vs 18.1 T scalar, scripts/
// ubench_f32x2.cu). The real-valued stream uses the scalar forms.
using Packed = unsigned long long; // {lo = re, hi = im}

GR4B200_HD Packed packPair(float lo, float hi) { return (static_cast<Packed>(__builtin_bit_cast(unsigned, hi)) << 32) | __builtin_bit_cast(unsigned, lo); }
GR4B200_HD float  packedLo(Packed v) { return __builtin_bit_cast(float, static_cast<unsigned>(v)); }
GR4B200_HD float  packedHi(Packed v) { return __builtin_bit_cast(float, static_cast<unsigned>(v >> 32)); }

// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (observed with nvcc 12.9, also under -fmad=false), which
// would silently turn the reference's two roundings into one. The packed exact forms are therefore written as two
// explicit FMAs against RUN-TIME constants the assembler cannot fold: fma(a, b, -0) is the correctly rounded product
// (adding -0 never changes a value or a zero's sign), fma(p, 1, acc) is the correctly rounded sum.
struct RoundingConsts {
    float one;     // 1.0f, read from the kernel arguments
    float negZero; // -0.0f, read from the kernel arguments
};

GR4B200_HD Packed mulV(float tap, Packed x, const RoundingConsts& k) {
#ifdef __CUDA_ARCH__
    Packed t, z, d;
    asm("mov.b64 %0, {%1, %1};" : "=l"(t) : "f"(tap));
    asm("mov.b64 %0, {%1, %1};" : "=l"(z) : "f"(k.negZero));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(t), "l"(x), "l"(z));
    return d;
#else
    (void)k;
    return packPair(tap * packedLo(x), tap * packedHi(x));
#endif
}
GR4B200_HD Packed addV(Packed a, Packed b, const RoundingConsts& k) {
#ifdef __CUDA_ARCH__
    Packed o, d;
    asm("mov.b64 %0, {%1, %1};" : "=l"(o) : "f"(k.one));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(b), "l"(o), "l"(a));
    return d;
#else
    (void)k;
    return packPair(packedLo(a) + packedLo(b), packedHi(a) + packedHi(b));
#endif
}
GR4B200_HD Packed fmaV(float tap, Packed x, Packed acc) {
#ifdef __CUDA_ARCH__
    Packed t, d;
    asm("mov.b64 %0, {%1, %1};" : "=l"(t) : "f"(tap));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(t), "l"(x), "l"(acc));
    return d;
#else
    return packPair(fmaf(tap, packedLo(x), packedLo(acc)), fmaf(tap, packedHi(x), packedHi(acc)));
#endif
}
GR4B200_HD float mulV(float tap, float x, const RoundingConsts&) { return fmulRn(tap, x); }
GR4B200_HD float addV(float a, float b, const RoundingConsts&) { return faddRn(a, b); }
GR4B200_HD float fmaV(float tap, float x, float acc) { return fmaRn(tap, x, acc); }

template<typename T>
struct VecOf;
template<>
struct VecOf<float> {
    using type = float;
    static GR4B200_HD float load(const float* p) { return *p; }
    static GR4B200_HD float store(float v) { return v; }
    static GR4B200_HD float zero() { return 0.f; }
    static GR4B200_HD float negZero() { return -0.f; }
};
template<>
struct VecOf<float2> {
    using type = Packed;
    static GR4B200_HD Packed load(const float2* p) { return *reinterpret_cast<const Packed*>(p); }
    static GR4B200_HD float2 store(Packed v) { return make_float2(packedLo(v), packedHi(v)); }
    static GR4B200_HD Packed zero() { return 0ull; }
    static GR4B200_HD Packed negZero() { return 0x8000000080000000ull; }
};
GR4B200_HD float  zeroOf(float) { return 0.f; }
GR4B200_HD float2 zeroOf(float2) { return make_float2(0.f, 0.f); }

// taps of lane j are stored contiguously: tapsT[j * lanePitch + m] = b[j + 16 m]
GR4B200_HD int lanePitchFor(int nTaps) { return ((nTaps + kLanes - 1) / kLanes + 7) / 8 * 8; }
// shared-memory bytes in front of the sample stages: natural-order taps + lane-major taps, rounded to 128 bytes
GR4B200_HD size_t tapsSmemBytes(int nTaps) { return (static_cast<size_t>((nTaps + 31) / 32 * 32 + kLanes * lanePitchFor(nTaps)) * sizeof(float) + 127) / 128 * 128; }
constexpr int kOutputsPerThreadD1 = 16; // outputs per thread of the full-rate kernel (tile = 256 threads * 16 = 4096 samples)

// MHere (1..8) consecutive taps of one lane against the sliding window of this thread's R outputs:
//   acc[r] (+)= b[j + 16 (mBase + m8)] * x[n0 + 16 r - j - 16 (mBase + m8)],   p[q] = x[n0 - j - 16 mBase + q]
// The window entry p[16 (i - 7)] serves every (r, m8) with r - m8 + 7 == i; walking i downwards visits each acc[r] in
// ascending m8, i.e. in the reference's accumulation order, while only one window entry is live at a time.
template<typename T, int R, int MHere, bool Exact>
GR4B200_HD void firLaneBlock(const T* p, const float* tapRow, typename VecOf<T>::type (&acc)[R], const RoundingConsts& k) {
    using V = VecOf<T>;
    float tap[MHere];
    if constexpr (MHere == 8) {
        const float4 lo = *reinterpret_cast<const float4*>(tapRow);
        const float4 hi = *reinterpret_cast<const float4*>(tapRow + 4);
        tap[0] = lo.x, tap[1] = lo.y, tap[2] = lo.z, tap[3] = lo.w, tap[4] = hi.x, tap[5] = hi.y, tap[6] = hi.z, tap[7] = hi.w;
    } else {
#pragma unroll
        for (int m8 = 0; m8 < MHere; ++m8) {
            tap[m8] = tapRow[m8];
        }
    }
#pragma unroll
    for (int i = R + 6; i >= 8 - MHere; --i) {
        const typename V::type w = V::load(p + kLanes * (i - 7));
#pragma unroll
        for (int m8 = 0; m8 < MHere; ++m8) {
            const int r = i - 7 + m8;
            if (r >= 0 && r < R) {
                if constexpr (Exact) {
                    acc[r] = addV(acc[r], mulV(tap[m8], w, k), k);
                } else {
                    acc[r] = fmaV(tap[m8], w, acc[r]);
                }
            }
        }
    }
}

// One thread's R outputs n0 + 16 r (full-rate indices): total[r] = sum_k b[k] x[n0 + 16 r - k] in the reference order.
// sBase[q] = x[n0 + q] for q in [-(nTaps-1), 16 (R-1)]; sTaps = the nTaps coefficients; sTapsT = lane-major copy.
template<typename T, int R, bool Exact>
GR4B200_HD void firThreadCompute(const T* sBase, const float* sTaps, const float* sTapsT, int nTaps, const RoundingConsts& k, T (&out)[R]) {
    using V   = VecOf<T>;
    using Vec = typename V::type;
    Vec total[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        total[r] = V::zero(); // the reference's init = T{0}
    }
    if (nTaps > 2 * kLanes) {
        const int fullBlocks = nTaps / kLanes; // every lane has at least this many taps
        const int remainder  = nTaps % kLanes;
        const int pitch      = lanePitchFor(nTaps);
#pragma unroll 1
        for (int j = 0; j < kLanes; ++j) {
            const int    mCount = fullBlocks + (j < remainder ? 1 : 0);
            const T*     p      = sBase - j;
            const float* tapRow = sTapsT + j * pitch;
            // Exact: lane[j] starts as the bare product f(j); (-0) + f == f bit for bit, so -0 is the neutral start.
            // Fast: accumulate straight into the output register.
            Vec acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                acc[r] = Exact ? V::negZero() : total[r];
            }
            int mBase = 0;
#pragma unroll 1
            for (; mBase + 8 <= mCount; mBase += 8) {
                firLaneBlock<T, R, 8, Exact>(p - kLanes * mBase, tapRow + mBase, acc, k);
            }
            const T*     pr = p - kLanes * mBase;
            const float* tr = tapRow + mBase;
            switch (mCount - mBase) { // uniform over the CTA
            case 1: firLaneBlock<T, R, 1, Exact>(pr, tr, acc, k); break;
            case 2: firLaneBlock<T, R, 2, Exact>(pr, tr, acc, k); break;
            case 3: firLaneBlock<T, R, 3, Exact>(pr, tr, acc, k); break;
            case 4: firLaneBlock<T, R, 4, Exact>(pr, tr, acc, k); break;
            case 5: firLaneBlock<T, R, 5, Exact>(pr, tr, acc, k); break;
            case 6: firLaneBlock<T, R, 6, Exact>(pr, tr, acc, k); break;
            case 7: firLaneBlock<T, R, 7, Exact>(pr, tr, acc, k); break;
            default: break;
            }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                total[r] = Exact ? addV(total[r], acc[r], k) : acc[r]; // init = init + lane[j]
            }
        }
    } else { // short filters: the reference folds left to right, init + f(0) + f(1) + ...
        for (int tapIndex = 0; tapIndex < nTaps; ++tapIndex) {
            const float tap = sTaps[tapIndex];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const Vec w = V::load(sBase + kLanes * r - tapIndex);
                total[r]    = Exact ? addV(total[r], mulV(tap, w, k), k) : fmaV(tap, w, total[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        out[r] = V::store(total[r]);
    }
}

// Threads: number of threads per CTA; R: outputs per thread; DLog2: log2(decimation), decimation | 16
template<typename T, int Threads, int R, int DLog2, bool Exact>
struct FirConfig {
    static constexpr int D        = 1 << DLog2;
    static constexpr int G        = kLanes / D;           // threads per 16-sample group
    static constexpr int Segments = Threads / G;          // groups of 16*R full-rate samples per tile
    static constexpr int TileIn   = Segments * kLanes * R; // full-rate samples per tile
};

// One thread of one tile: sTile holds x[tileStart - haloPad .. tileStart + TileIn), outputs go to out[(tileStart + n)/D].
template<typename T, int Threads, int R, int DLog2, bool Exact>
GR4B200_HD void firTileThread(int tid, const T* sTile, const float* sTaps, const float* sTapsT, int nTaps, int haloPad, long long tileStart, long long nOut, const RoundingConsts& k, T* out) {
    using Cfg       = FirConfig<T, Threads, R, DLog2, Exact>;
    const int seg   = tid / Cfg::G;
    const int tsub  = tid % Cfg::G;
    const int n0    = seg * (kLanes * R) + tsub * Cfg::D; // tile-relative full-rate index of this thread's first output
    const T*  sBase = sTile + haloPad + n0;               // sBase[q] = x[tileStart + n0 + q]
    T         total[R];
    firThreadCompute<T, R, Exact>(sBase, sTaps, sTapsT, nTaps, k, total);
    const long long outBase = (tileStart + n0) >> DLog2;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long o = outBase + ((kLanes * r) >> DLog2);
        if (o < nOut) {
            out[o] = total[r];
        }
    }
}

} // namespace gr4b200
