// Block-mode epilogue of the FFT kernels (device only): magnitude / phase of the FFT block
// (algorithm/include/gnuradio-4.0/algorithm/fourier/fft_common.hpp:22-123) and the four-plane DataSet store, shared by
// fftRadixKernel (fft.cu) and the fused FIR -> FFT kernel (fir_fft.cu).
#pragma once

#include "common.cuh"
#include "fft_radix.cuh"

namespace gr4b200 {

// fft_common.hpp:37-44: magnitude = hypot(re, im) * 2 / N (optionally 20 log10, log(0) -> lowest());
// fft_common.hpp:107:   phase     = atan2(im, re) (optionally degrees).
// Both come from one octant reduction, branch free: hi = max(|re|,|im|), t = min/hi in [0, 1];
//   hypot = hi * sqrt(1 + t^2)           (no overflow / underflow anywhere in the float range; <= 3 ulp)
//   atan  = odd minimax polynomial of degree 17 in t (Abramowitz & Stegun 4.4.49, |rel. error| <= 2e-8), folded back
//           through the octant; absolute error <= 3e-7 rad. The sign of a zero real part is honoured (signbit), so
//           atan2(+-0, -0) = +-pi and atan2(+-0, +0) = +-0 as the library has it. (inf, inf) gives NaN (library: pi/4).
// N is a power of two, so (m * 2) / N == m * (2 / N) bit for bit.
__device__ __forceinline__ float rcpApprox(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrtApprox(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
static __device__ __noinline__ float decibel(float mag) { return mag > 0.f ? __fmul_rn(20.f, log10f(mag)) : -3.402823466e+38f; }
__device__ __forceinline__ float toDegrees(float phase) { return __fmul_rn(__fmul_rn(phase, 180.f), 0.318309886183790671538f); }

__device__ __forceinline__ Cx splat(float v) { return cxMake(v, v); }

// two bins at a time: the polynomial runs on packed pairs
__device__ __forceinline__ void magnitudePhase2(Cx binA, Cx binB, float twoOverN, float& magA, float& magB, float& phA, float& phB) {
    float ra, ia, rb, ib;
    cxSplit(binA, ra, ia);
    cxSplit(binB, rb, ib);
    const float axa = fabsf(ra), aya = fabsf(ia), axb = fabsf(rb), ayb = fabsf(ib);
    const float hia = fmaxf(axa, aya), loa = fminf(axa, aya), hib = fmaxf(axb, ayb), lob = fminf(axb, ayb);
    const float ta  = loa * rcpApprox(fmaxf(hia, 1.17549435e-38f));
    const float tb  = lob * rcpApprox(fmaxf(hib, 1.17549435e-38f));
    const Cx    t   = cxMake(ta, tb);
    const Cx    t2  = pkMul(t, t);
    Cx          p   = pkFma(splat(0.0028662257f), t2, splat(-0.0161657367f));
    p               = pkFma(p, t2, splat(0.0429096138f));
    p               = pkFma(p, t2, splat(-0.0752896400f));
    p               = pkFma(p, t2, splat(0.1065626393f));
    p               = pkFma(p, t2, splat(-0.1420889944f));
    p               = pkFma(p, t2, splat(0.1999355085f));
    p               = pkFma(p, t2, splat(-0.3333314528f));
    p               = pkFma(pkMul(p, t2), t, t);
    const Cx h2     = pkAdd(t2, splat(1.f));
    magA            = (hia * twoOverN) * sqrtApprox(cxRe(h2));
    magB            = (hib * twoOverN) * sqrtApprox(cxIm(h2));
    float pa = cxRe(p), pb = cxIm(p);
    pa  = aya > axa ? 1.57079632679489661923f - pa : pa;
    pb  = ayb > axb ? 1.57079632679489661923f - pb : pb;
    pa  = __float_as_int(ra) < 0 ? 3.14159265358979323846f - pa : pa;
    pb  = __float_as_int(rb) < 0 ? 3.14159265358979323846f - pb : pb;
    phA = copysignf(pa, ia);
    phB = copysignf(pb, ib);
}

// The spectrum of one transform is parked in shared memory in natural order (fftPark); thread t of its T threads finishes
// four groups of four consecutive bins and writes 16-byte vectors into the four planes sig[4][N] =
// {magnitude (fft-shifted), phase (fft-shifted), Re, Im}; lo/hi collect the per-signal ranges when asked.
template<int N>
__device__ __forceinline__ void fftBlockEpilogue(int t, const Cx* park, float* __restrict__ sig, bool dB, bool deg, bool wantRanges, float (&lo)[4], float (&hi)[4], bool storeIt) {
    constexpr int T = FftGeom<N>::kThreads;
    const Cx* parkedLo = park + fftParkReadBase(t, 0);
    const Cx* parkedHi = park + fftParkReadBase(t, 1);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int  k0      = 4 * (g * T + t);
        const int  shifted = k0 ^ (N / 2); // fft-shift: (k0 + N/2) mod N, keeps groups of four together
        const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(parkedLo + 4 * g * T);
        const ulonglong2 b = *reinterpret_cast<const ulonglong2*>(parkedHi + 4 * g * T);
        const Cx    bins[4] = {a.x, a.y, b.x, b.y};
        float       mag[4], ph[4], re[4], im[4];
        magnitudePhase2(bins[0], bins[1], 2.f / N, mag[0], mag[1], ph[0], ph[1]);
        magnitudePhase2(bins[2], bins[3], 2.f / N, mag[2], mag[3], ph[2], ph[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            cxSplit(bins[e], re[e], im[e]);
        }
        if (dB || deg) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                mag[e] = dB ? decibel(mag[e]) : mag[e];
                ph[e]  = deg ? toDegrees(ph[e]) : ph[e];
            }
        }
        if (wantRanges) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                lo[0] = fminf(lo[0], mag[e]), hi[0] = fmaxf(hi[0], mag[e]);
                lo[1] = fminf(lo[1], ph[e]), hi[1] = fmaxf(hi[1], ph[e]);
                lo[2] = fminf(lo[2], re[e]), hi[2] = fmaxf(hi[2], re[e]);
                lo[3] = fminf(lo[3], im[e]), hi[3] = fmaxf(hi[3], im[e]);
            }
        }
        if (storeIt) {
            stStream4(reinterpret_cast<float4*>(sig + shifted), make_float4(mag[0], mag[1], mag[2], mag[3]));
            stStream4(reinterpret_cast<float4*>(sig + N + shifted), make_float4(ph[0], ph[1], ph[2], ph[3]));
            stStream4(reinterpret_cast<float4*>(sig + 2 * N + k0), make_float4(re[0], re[1], re[2], re[3]));
            stStream4(reinterpret_cast<float4*>(sig + 3 * N + k0), make_float4(im[0], im[1], im[2], im[3]));
        }
    }
}

// The same for the FFT block on a REAL stream (planes of H = N/2 values): magnitude and phase of bins [0, H) in natural
// order (no fft-shift for a half spectrum, fft_common.hpp:52,118), Re and Im of bins [H, N) -- what createDataset copies
// with `std::span{_outData}.last(N)` (fft.hpp:212-217). Thread t finishes two groups of four consecutive values per plane.
template<int N>
__device__ __forceinline__ void fftBlockEpilogueReal(int t, const Cx* park, float* __restrict__ sig, bool dB, bool deg, bool wantRanges, float (&lo)[4], float (&hi)[4], bool storeIt) {
    constexpr int T = FftGeom<N>::kThreads;
    constexpr int H = N / 2;
    const Cx* parkedLo = park + fftParkReadBase(t, 0);
    const Cx* parkedHi = park + fftParkReadBase(t, 1);
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        const int        k0 = 4 * (g * T + t);
        const ulonglong2 a  = *reinterpret_cast<const ulonglong2*>(parkedLo + 4 * g * T);
        const ulonglong2 b  = *reinterpret_cast<const ulonglong2*>(parkedHi + 4 * g * T);
        const ulonglong2 c  = *reinterpret_cast<const ulonglong2*>(parkedLo + 4 * (g + 2) * T); // bins H + k0 ..
        const ulonglong2 d  = *reinterpret_cast<const ulonglong2*>(parkedHi + 4 * (g + 2) * T);
        const Cx         upper[4] = {c.x, c.y, d.x, d.y};
        float            mag[4], ph[4], re[4], im[4];
        magnitudePhase2(a.x, a.y, 2.f / N, mag[0], mag[1], ph[0], ph[1]);
        magnitudePhase2(b.x, b.y, 2.f / N, mag[2], mag[3], ph[2], ph[3]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            cxSplit(upper[e], re[e], im[e]);
        }
        if (dB || deg) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                mag[e] = dB ? decibel(mag[e]) : mag[e];
                ph[e]  = deg ? toDegrees(ph[e]) : ph[e];
            }
        }
        if (wantRanges) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                lo[0] = fminf(lo[0], mag[e]), hi[0] = fmaxf(hi[0], mag[e]);
                lo[1] = fminf(lo[1], ph[e]), hi[1] = fmaxf(hi[1], ph[e]);
                lo[2] = fminf(lo[2], re[e]), hi[2] = fmaxf(hi[2], re[e]);
                lo[3] = fminf(lo[3], im[e]), hi[3] = fmaxf(hi[3], im[e]);
            }
        }
        if (storeIt) {
            stStream4(reinterpret_cast<float4*>(sig + k0), make_float4(mag[0], mag[1], mag[2], mag[3]));
            stStream4(reinterpret_cast<float4*>(sig + H + k0), make_float4(ph[0], ph[1], ph[2], ph[3]));
            stStream4(reinterpret_cast<float4*>(sig + 2 * H + k0), make_float4(re[0], re[1], re[2], re[3]));
            stStream4(reinterpret_cast<float4*>(sig + 3 * H + k0), make_float4(im[0], im[1], im[2], im[3]));
        }
    }
}

} // namespace gr4b200
