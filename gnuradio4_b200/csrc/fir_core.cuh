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

// ---- element helpers: T = float (real stream) or float2 (complex stream: re/im filtered independently) -----------
template<bool Exact>
GR4B200_HD float mulTap(float b, float x) {
    return fmulRn(b, x);
}
template<bool Exact>
GR4B200_HD float2 mulTap(float b, float2 x) {
    return make_float2(fmulRn(b, x.x), fmulRn(b, x.y));
}
// acc (+)= b*x: Exact => round the product, then the sum; Fast => one fused operation
template<bool Exact>
GR4B200_HD float macTap(float acc, float b, float x) {
    if constexpr (Exact) {
        return faddRn(acc, fmulRn(b, x));
    } else {
        return fmaRn(b, x, acc);
    }
}
template<bool Exact>
GR4B200_HD float2 macTap(float2 acc, float b, float2 x) {
    return make_float2(macTap<Exact>(acc.x, b, x.x), macTap<Exact>(acc.y, b, x.y));
}
GR4B200_HD float  addRn(float a, float b) { return faddRn(a, b); }
GR4B200_HD float2 addRn(float2 a, float2 b) { return make_float2(faddRn(a.x, b.x), faddRn(a.y, b.y)); }
GR4B200_HD float  zeroOf(float) { return 0.f; }
GR4B200_HD float2 zeroOf(float2) { return make_float2(0.f, 0.f); }


// One thread's R outputs n0 + 16 r (full-rate indices): total[r] = sum_k b[k] x[n0 + 16 r - k] in the reference order.
// sBase[q] = x[n0 + q] for q in [-(nTaps-1), 16 (R-1)]; sTaps = the nTaps coefficients.
template<typename T, int R, bool Exact>
GR4B200_HD void firThreadCompute(const T* sBase, const float* sTaps, int nTaps, T (&total)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        total[r] = zeroOf(T{});
    }

    if (nTaps > 2 * kLanes) {
        const int fullBlocks = nTaps / kLanes; // every lane has at least this many taps
        const int remainder  = nTaps % kLanes;
        for (int j = 0; j < kLanes; ++j) {
            const int mCount = fullBlocks + (j < remainder ? 1 : 0);
            T         acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                acc[r] = zeroOf(T{});
            }
            for (int mBase = 0; mBase < mCount; mBase += 8) {
                const int mHere = mCount - mBase < 8 ? mCount - mBase : 8; // taps of this lane in this block (uniform)
                // window[i] = x[n0 - j - 16*mBase + 16*(i-7)], entries below 7-(mHere-1) are not needed
                T         window[R + 7];
                const T*  p = sBase - j - kLanes * mBase;
#pragma unroll
                for (int i = 0; i < R + 7; ++i) {
                    if (i >= 8 - mHere) {
                        window[i] = p[kLanes * (i - 7)];
                    } else {
                        window[i] = zeroOf(T{});
                    }
                }
#pragma unroll
                for (int m8 = 0; m8 < 8; ++m8) {
                    if (m8 < mHere) {
                        const float tap = sTaps[j + kLanes * (mBase + m8)];
                        if (Exact && mBase + m8 == 0) { // lane[j] starts as the bare product f(j)
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                acc[r] = mulTap<Exact>(tap, window[r - m8 + 7]);
                            }
                        } else if (Exact) {
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                acc[r] = macTap<true>(acc[r], tap, window[r - m8 + 7]);
                            }
                        } else {
#pragma unroll
                            for (int r = 0; r < R; ++r) {
                                total[r] = macTap<false>(total[r], tap, window[r - m8 + 7]);
                            }
                        }
                    }
                }
            }
            if constexpr (Exact) { // fold lane j: init = init + lane[j]
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    total[r] = addRn(total[r], acc[r]);
                }
            }
        }
    } else { // short filters: the reference folds left to right, init + f(0) + f(1) + ...
        for (int k = 0; k < nTaps; ++k) {
            const float tap = sTaps[k];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                total[r] = macTap<Exact>(total[r], tap, sBase[kLanes * r - k]);
            }
        }
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
GR4B200_HD void firTileThread(int tid, const T* sTile, const float* sTaps, int nTaps, int haloPad, long long tileStart, long long nOut, T* out) {
    using Cfg       = FirConfig<T, Threads, R, DLog2, Exact>;
    const int seg   = tid / Cfg::G;
    const int tsub  = tid % Cfg::G;
    const int n0    = seg * (kLanes * R) + tsub * Cfg::D; // tile-relative full-rate index of this thread's first output
    const T*  sBase = sTile + haloPad + n0;               // sBase[q] = x[tileStart + n0 + q]
    T         total[R];
    firThreadCompute<T, R, Exact>(sBase, sTaps, nTaps, total);
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
