// Register-level building blocks of the power-of-two FFT family (N = 16 .. 8192), shared by the device kernels
// (fft.cu) and the host emulation of the CPU-side tests (tests/host_emulation.cu).
//
// Contract (algorithm/include/gnuradio-4.0/algorithm/fourier/fft.hpp:113-153): unnormalised forward DFT,
// X[k] = sum_n x[n] exp(-j 2 pi k n / N), natural order in and out, float arithmetic.
//
// Formulation: Stockham autosort, decimation in time, radices {16, 16, ..., r} with r = N / 16^p in {2, 4, 8, 16}.
// A transform is owned by T = N/16 threads; in EVERY pass thread t holds the 16 points  v[m] = data[t + T*m]  (m = 0..15)
// in registers, so all reads are lane-contiguous. Pass p has Ns = 16^p, radix R and G = 16/R butterflies per thread:
//   butterfly g works on j = t + T*g:   inputs  v[g + G*q],  q = 0..R-1,   twiddled by W_{Ns*R}^((j mod Ns) * q),
//   output q lands at index (j / Ns) * Ns * R + (j mod Ns) + q * Ns of the next pass's array.
// In the last pass that index is t + T*(g + G*q): register m again holds X[t + T*m] -- coalesced stores.
// Exchange arrays live in shared memory at pad(i) = i + i/16: every 8-byte access pattern above is bank-conflict free.
//
// A complex value is one packed f32x2 register pair (Cx): sm_100a's FADD2 / FMUL2 / FFMA2 take one issue slot for both
// halves, support a scalar-broadcast operand and a swap-and-negate-one-half operand modifier (which is exactly a
// multiplication by +-j), so a radix-4 butterfly is 8 instructions and a complex product is 2.
#pragma once

#include <cuda_runtime.h>

#include <cmath>
#include <cstddef>
#include <cstring>

#ifndef GR4B200_HD
#define GR4B200_HD __host__ __device__ __forceinline__
#endif

namespace gr4b200 {

using Cx = unsigned long long; // {low 32 bits = re, high 32 bits = im}: the memory layout of std::complex<float>

GR4B200_HD Cx cxMake(float re, float im) {
#ifdef __CUDA_ARCH__
    Cx r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(re), "f"(im));
    return r;
#else
    unsigned lo, hi;
    std::memcpy(&lo, &re, 4);
    std::memcpy(&hi, &im, 4);
    return (static_cast<Cx>(hi) << 32) | lo;
#endif
}
GR4B200_HD void cxSplit(Cx v, float& re, float& im) {
#ifdef __CUDA_ARCH__
    asm("mov.b64 {%0, %1}, %2;" : "=f"(re), "=f"(im) : "l"(v));
#else
    const unsigned lo = static_cast<unsigned>(v), hi = static_cast<unsigned>(v >> 32);
    std::memcpy(&re, &lo, 4);
    std::memcpy(&im, &hi, 4);
#endif
}
GR4B200_HD float cxRe(Cx v) {
    float re, im;
    cxSplit(v, re, im);
    return re;
}
GR4B200_HD float cxIm(Cx v) {
    float re, im;
    cxSplit(v, re, im);
    return im;
}

// packed f32x2 arithmetic (device only)
__device__ __forceinline__ Cx pkAdd(Cx a, Cx b) {
    Cx r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ Cx pkSub(Cx a, Cx b) {
    Cx r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ Cx pkMul(Cx a, Cx b) {
    Cx r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ Cx pkFma(Cx a, Cx b, Cx c) {
    Cx r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ Cx pkSwap(Cx v) { // (im, re): folded into the consumer's operand modifier by ptxas
    float re, im;
    cxSplit(v, re, im);
    return cxMake(im, re);
}

GR4B200_HD Cx cxAdd(Cx a, Cx b) {
#ifdef __CUDA_ARCH__
    return pkAdd(a, b);
#else
    return cxMake(cxRe(a) + cxRe(b), cxIm(a) + cxIm(b));
#endif
}
GR4B200_HD Cx cxSub(Cx a, Cx b) {
#ifdef __CUDA_ARCH__
    return pkSub(a, b);
#else
    return cxMake(cxRe(a) - cxRe(b), cxIm(a) - cxIm(b));
#endif
}
// a - j b = (a.re + b.im, a.im - b.re)
GR4B200_HD Cx cxAddMinusJ(Cx a, Cx b) {
#ifdef __CUDA_ARCH__
    return pkFma(pkSwap(b), cxMake(1.f, -1.f), a);
#else
    return cxMake(cxRe(a) + cxIm(b), cxIm(a) - cxRe(b));
#endif
}
// a + j b = (a.re - b.im, a.im + b.re)
GR4B200_HD Cx cxAddPlusJ(Cx a, Cx b) {
#ifdef __CUDA_ARCH__
    return pkFma(pkSwap(b), cxMake(-1.f, 1.f), a);
#else
    return cxMake(cxRe(a) - cxIm(b), cxIm(a) + cxRe(b));
#endif
}
// -j a = (a.im, -a.re)
GR4B200_HD Cx cxMulMinusJ(Cx a) {
#ifdef __CUDA_ARCH__
    return pkMul(pkSwap(a), cxMake(1.f, -1.f));
#else
    return cxMake(cxIm(a), -cxRe(a));
#endif
}
GR4B200_HD Cx cxScale(Cx a, float s) {
#ifdef __CUDA_ARCH__
    return pkMul(a, cxMake(s, s));
#else
    return cxMake(cxRe(a) * s, cxIm(a) * s);
#endif
}
// a * w = w.re * a + w.im * (j a)
GR4B200_HD Cx cxMul(Cx a, Cx w) {
#ifdef __CUDA_ARCH__
    const float wr = cxRe(w), wi = cxIm(w);
    return pkFma(cxMake(wr, wr), a, pkMul(cxMake(wi, wi), pkMul(pkSwap(a), cxMake(-1.f, 1.f))));
#else
    const float ar = cxRe(a), ai = cxIm(a), wr = cxRe(w), wi = cxIm(w);
    return cxMake(fmaf(wr, ar, -(wi * ai)), fmaf(wr, ai, wi * ar));
#endif
}
// a * (c - j s) = c * a + s * (-j a), c and s compile-time constants
GR4B200_HD Cx cxMulConst(Cx a, float c, float s) {
#ifdef __CUDA_ARCH__
    return pkFma(cxMake(c, c), a, pkMul(cxMake(s, s), pkMul(pkSwap(a), cxMake(1.f, -1.f))));
#else
    const float ar = cxRe(a), ai = cxIm(a);
    return cxMake(fmaf(c, ar, s * ai), fmaf(c, ai, -(s * ar)));
#endif
}

constexpr float kRootHalf = 0.70710678118654752440f;
constexpr float kCosPi8_  = 0.92387953251128675613f;
constexpr float kSinPi8_  = 0.38268343236508977173f;

// ---- small DFTs, natural order in and out ---------------------------------------------------------------------------
GR4B200_HD void cxDft2(Cx& x0, Cx& x1) {
    const Cx s = cxAdd(x0, x1), d = cxSub(x0, x1);
    x0 = s;
    x1 = d;
}
GR4B200_HD void cxDft4(Cx& x0, Cx& x1, Cx& x2, Cx& x3) {
    const Cx s02 = cxAdd(x0, x2), d02 = cxSub(x0, x2), s13 = cxAdd(x1, x3), d13 = cxSub(x1, x3);
    x0 = cxAdd(s02, s13);
    x2 = cxSub(s02, s13);
    x1 = cxAddMinusJ(d02, d13);
    x3 = cxAddPlusJ(d02, d13);
}
GR4B200_HD void cxDft8(Cx& x0, Cx& x1, Cx& x2, Cx& x3, Cx& x4, Cx& x5, Cx& x6, Cx& x7) {
    Cx a0 = cxAdd(x0, x4), b0 = cxSub(x0, x4);
    Cx a1 = cxAdd(x1, x5), b1 = cxSub(x1, x5);
    Cx a2 = cxAdd(x2, x6), b2 = cxSub(x2, x6);
    Cx a3 = cxAdd(x3, x7), b3 = cxSub(x3, x7);
    b1    = cxScale(cxAddMinusJ(b1, b1), kRootHalf);  // * W8^1 = (1 - j) / sqrt 2
    b2    = cxMulMinusJ(b2);                          // * W8^2 = -j
    b3    = cxScale(cxAddPlusJ(b3, b3), -kRootHalf);  // * W8^3 = -(1 + j) / sqrt 2
    cxDft4(a0, a1, a2, a3);                           // X[0], X[2], X[4], X[6]
    cxDft4(b0, b1, b2, b3);                           // X[1], X[3], X[5], X[7]
    x0 = a0, x1 = b0, x2 = a1, x3 = b1, x4 = a2, x5 = b2, x6 = a3, x7 = b3;
}
// n = 4a + b, k = c + 4d: y[b][c] = sum_a x[4a+b] W4^(ac);  y[b][c] *= W16^(bc);  X[c+4d] = sum_b y[b][c] W4^(bd)
GR4B200_HD void cxDft16(Cx (&x)[16]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        cxDft4(x[b], x[4 + b], x[8 + b], x[12 + b]); // x[4c + b] = y[b][c]
    }
    x[5]  = cxMulConst(x[5], kCosPi8_, kSinPi8_);             // W16^1
    x[6]  = cxScale(cxAddMinusJ(x[6], x[6]), kRootHalf);      // W16^2
    x[7]  = cxMulConst(x[7], kSinPi8_, kCosPi8_);             // W16^3
    x[9]  = cxScale(cxAddMinusJ(x[9], x[9]), kRootHalf);      // W16^2
    x[10] = cxMulMinusJ(x[10]);                               // W16^4
    x[11] = cxScale(cxAddPlusJ(x[11], x[11]), -kRootHalf);    // W16^6
    x[13] = cxMulConst(x[13], kSinPi8_, kCosPi8_);            // W16^3
    x[14] = cxScale(cxAddPlusJ(x[14], x[14]), -kRootHalf);    // W16^6
    x[15] = cxMulConst(x[15], -kCosPi8_, -kSinPi8_);          // W16^9
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        cxDft4(x[4 * c + 0], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]); // x[4c + d] = X[c + 4d]
    }
    Cx t;
#define GR4B200_CXSWAP(i, j) t = x[i], x[i] = x[j], x[j] = t;
    GR4B200_CXSWAP(1, 4)
    GR4B200_CXSWAP(2, 8)
    GR4B200_CXSWAP(3, 12)
    GR4B200_CXSWAP(6, 9)
    GR4B200_CXSWAP(7, 13)
    GR4B200_CXSWAP(11, 14)
#undef GR4B200_CXSWAP
}

// ---- geometry ----------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int fftPad(int i) { return i + (i >> 4); }

template<int N>
struct FftGeom {
    static_assert(N >= 16 && N <= 8192 && (N & (N - 1)) == 0, "N must be a power of two in [16, 8192]");
    static constexpr int kThreads = N / 16;                                    // T: threads per transform
    static constexpr int kPasses  = N == 16 ? 1 : (N <= 256 ? 2 : (N <= 4096 ? 3 : 4));
    static constexpr int kCta     = kThreads > 256 ? kThreads : 256;           // threads per CTA
    static constexpr int kPerCta  = kCta / kThreads;                           // transforms per CTA iteration
    static constexpr int kPadded  = N + N / 16;                                // elements of one padded exchange array
    __host__ __device__ static constexpr int ns(int p) { return p == 0 ? 1 : (p == 1 ? 16 : (p == 2 ? 256 : 4096)); }
    __host__ __device__ static constexpr int radix(int p) { return N / ns(p) >= 16 ? 16 : N / ns(p); }
    __host__ __device__ static constexpr int log2Radix(int p) { return radix(p) == 16 ? 4 : (radix(p) == 8 ? 3 : (radix(p) == 4 ? 2 : 1)); }
    // twiddle table of pass p >= 1: [log2 R][Ns] entries W_{Ns R}^(2^i e)
    __host__ __device__ static constexpr int tableSize(int p) { return p == 0 ? 0 : log2Radix(p) * ns(p); }
    __host__ __device__ static constexpr int tableOffset(int p) { return p <= 1 ? 0 : tableOffset(p - 1) + tableSize(p - 1); }
    static constexpr int kTableEntries = tableOffset(kPasses - 1) + tableSize(kPasses - 1);
};

// host: the twiddle tables of all passes, computed in double and rounded once
template<int N>
inline void fftFillTables(float2* table) {
    using G = FftGeom<N>;
    for (int p = 1; p < G::kPasses; ++p) {
        const int    ns = G::ns(p), r = G::radix(p);
        const double span = static_cast<double>(ns) * r;
        for (int i = 0; i < G::log2Radix(p); ++i) {
            for (int e = 0; e < ns; ++e) {
                const double arg = -2.0 * 3.14159265358979323846 * static_cast<double>((static_cast<long long>(e) << i) % (static_cast<long long>(ns) * r)) / span;
                table[G::tableOffset(p) + i * ns + e] = make_float2(static_cast<float>(cos(arg)), static_cast<float>(sin(arg)));
            }
        }
    }
}

GR4B200_HD Cx cxLoadTable(const float2* p) {
#ifdef __CUDA_ARCH__
    const float2 v = __ldg(p);
    return cxMake(v.x, v.y);
#else
    return cxMake(p->x, p->y);
#endif
}

// x[k] *= w^k, k = 1..15, from w^1, w^2, w^4, w^8 (table values): at most three complex products per power
GR4B200_HD void cxApplyPowers16(Cx (&x)[16], Cx w1, Cx w2, Cx w4, Cx w8) {
    const Cx w3 = cxMul(w2, w1), w5 = cxMul(w4, w1), w6 = cxMul(w4, w2), w7 = cxMul(w4, w3);
    x[1]  = cxMul(x[1], w1);
    x[2]  = cxMul(x[2], w2);
    x[3]  = cxMul(x[3], w3);
    x[4]  = cxMul(x[4], w4);
    x[5]  = cxMul(x[5], w5);
    x[6]  = cxMul(x[6], w6);
    x[7]  = cxMul(x[7], w7);
    x[8]  = cxMul(x[8], w8);
    x[9]  = cxMul(x[9], cxMul(w8, w1));
    x[10] = cxMul(x[10], cxMul(w8, w2));
    x[11] = cxMul(x[11], cxMul(w8, w3));
    x[12] = cxMul(x[12], cxMul(w8, w4));
    x[13] = cxMul(x[13], cxMul(w8, w5));
    x[14] = cxMul(x[14], cxMul(w8, w6));
    x[15] = cxMul(x[15], cxMul(w8, w7));
}

// ---- one pass on the registers of thread t: twiddles (p > 0) and butterflies ---------------------------------------
// The table entries a thread needs depend on t only, never on the transform: a persistent thread can load them once
// (fftLoadTwiddles) and reuse them for every transform it processes (fftPassWithTwiddles).
constexpr int kFftTwiddleRegs = 8; // radix 16: W^e, W^2e, W^4e, W^8e; radix r < 16: log2(r) entries for each of the 16/r butterflies

template<int N, int P>
GR4B200_HD void fftLoadTwiddles(int t, const float2* tables, Cx (&tw)[kFftTwiddleRegs]) {
    using G              = FftGeom<N>;
    constexpr int R      = G::radix(P);
    constexpr int Ns     = G::ns(P);
    constexpr int Groups = 16 / R;
    constexpr int T      = G::kThreads;
    if constexpr (P > 0) {
        const float2* table = tables + G::tableOffset(P);
        if constexpr (R == 16) {
            const int e = t & (Ns - 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                tw[i] = cxLoadTable(table + i * Ns + e);
            }
        } else {
            // last pass: Ns * R = N, butterfly g works on j = t + T g < Ns
#pragma unroll
            for (int g = 0; g < Groups; ++g) {
#pragma unroll
                for (int i = 0; i < G::log2Radix(P); ++i) {
                    tw[g * G::log2Radix(P) + i] = cxLoadTable(table + i * Ns + t + T * g);
                }
            }
        }
    }
}

template<int N, int P>
GR4B200_HD void fftPassWithTwiddles(Cx (&v)[16], const Cx (&tw)[kFftTwiddleRegs]) {
    using G              = FftGeom<N>;
    constexpr int R      = G::radix(P);
    constexpr int Groups = 16 / R;
    if constexpr (R == 16) {
        if constexpr (P > 0) {
            cxApplyPowers16(v, tw[0], tw[1], tw[2], tw[3]);
        }
        cxDft16(v);
    } else {
        static_assert(P > 0, "the first pass is always radix 16");
#pragma unroll
        for (int g = 0; g < Groups; ++g) {
            if constexpr (R == 2) {
                v[g + Groups] = cxMul(v[g + Groups], tw[g]);
                cxDft2(v[g], v[g + Groups]);
            } else if constexpr (R == 4) {
                const Cx w1 = tw[2 * g], w2 = tw[2 * g + 1];
                v[g + Groups]     = cxMul(v[g + Groups], w1);
                v[g + 2 * Groups] = cxMul(v[g + 2 * Groups], w2);
                v[g + 3 * Groups] = cxMul(v[g + 3 * Groups], cxMul(w2, w1));
                cxDft4(v[g], v[g + Groups], v[g + 2 * Groups], v[g + 3 * Groups]);
            } else {
                const Cx w1 = tw[3 * g], w2 = tw[3 * g + 1], w4 = tw[3 * g + 2];
                const Cx w3 = cxMul(w2, w1);
                v[g + Groups]     = cxMul(v[g + Groups], w1);
                v[g + 2 * Groups] = cxMul(v[g + 2 * Groups], w2);
                v[g + 3 * Groups] = cxMul(v[g + 3 * Groups], w3);
                v[g + 4 * Groups] = cxMul(v[g + 4 * Groups], w4);
                v[g + 5 * Groups] = cxMul(v[g + 5 * Groups], cxMul(w4, w1));
                v[g + 6 * Groups] = cxMul(v[g + 6 * Groups], cxMul(w4, w2));
                v[g + 7 * Groups] = cxMul(v[g + 7 * Groups], cxMul(w4, w3));
                cxDft8(v[g], v[g + Groups], v[g + 2 * Groups], v[g + 3 * Groups], v[g + 4 * Groups], v[g + 5 * Groups], v[g + 6 * Groups], v[g + 7 * Groups]);
            }
        }
    }
}

template<int N, int P>
GR4B200_HD void fftPassCompute(int t, Cx (&v)[16], const float2* tables) {
    Cx tw[kFftTwiddleRegs];
    fftLoadTwiddles<N, P>(t, tables, tw);
    fftPassWithTwiddles<N, P>(v, tw);
}

// index (unpadded) in the next pass's array of register m after a NON-final pass P (always radix 16)
template<int N, int P>
GR4B200_HD int fftScatterIndex(int t, int m) {
    constexpr int Ns = FftGeom<N>::ns(P);
    return (t / Ns) * (16 * Ns) + (t & (Ns - 1)) + m * Ns;
}

// registers <-> exchange array at pad(i) = i + i/16, thread t of the transform. Both index families are affine in m
// once the thread's base is known (the per-register part is a compile-time offset):
//   scatter, Ns = 1 : pad(16 t + m)                 = 17 t + m
//   scatter, Ns >= 16: pad(b + m Ns), b = (t/Ns) 16 Ns + t%Ns  = pad(b) + m (Ns + Ns/16)   (b + m Ns keeps b's low four bits)
//   gather, T >= 16  : pad(t + T m)                 = pad(t) + m (T + T/16)
//   gather, T < 16   : pad(t + T m)                 = t + (T m + T m / 16)                  (t < T, T | 16)
template<int N, int P>
GR4B200_HD void fftScatter(int t, const Cx (&v)[16], Cx* array) {
    constexpr int Ns = FftGeom<N>::ns(P);
    if constexpr (Ns == 1) {
        Cx* base = array + 17 * t;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            base[m] = v[m];
        }
    } else {
        const int b    = (t / Ns) * (16 * Ns) + (t & (Ns - 1));
        Cx*       base = array + b + (b >> 4);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            base[m * (Ns + Ns / 16)] = v[m];
        }
    }
}
template<int N>
GR4B200_HD void fftGather(int t, const Cx* array, Cx (&v)[16]) {
    constexpr int T = FftGeom<N>::kThreads;
    if constexpr (T >= 16) {
        const Cx* base = array + t + (t >> 4);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            v[m] = base[m * (T + T / 16)];
        }
    } else {
        const Cx* base = array + t;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            v[m] = base[T * m + ((T * m) >> 4)];
        }
    }
}

// window in the per-thread layout windowT[16 t + m] = w[t + T m] (four 16-byte loads per thread, again a function of t
// only); blocks/fourier/include/gnuradio-4.0/fourier/fft.hpp:155-162 multiplies re and im by w[n] before the transform
GR4B200_HD void fftLoadWindow(int t, const float* windowT, float (&w)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#ifdef __CUDA_ARCH__
        const float4 f = __ldg(reinterpret_cast<const float4*>(windowT + 16 * t) + q);
#else
        const float4 f = reinterpret_cast<const float4*>(windowT + 16 * t)[q];
#endif
        w[4 * q + 0] = f.x;
        w[4 * q + 1] = f.y;
        w[4 * q + 2] = f.z;
        w[4 * q + 3] = f.w;
    }
}
GR4B200_HD void fftApplyWindow(const float (&w)[16], Cx (&v)[16]) {
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        v[m] = cxScale(v[m], w[m]);
    }
}
GR4B200_HD void fftApplyWindow(int t, const float* windowT, Cx (&v)[16]) {
    float w[16];
    fftLoadWindow(t, windowT, w);
    fftApplyWindow(w, v);
}

// natural-order parking slot of bin k for the block-mode epilogue: 16-byte reads of four consecutive bins by
// consecutive lanes (stride 32 bytes) would be 2-way bank conflicted; swapping the two 16-byte halves of every other
// 64-byte group makes both the 8-byte writes (k = t + T m) and the 16-byte reads conflict free
GR4B200_HD int fftParkSlot(int k) { return k ^ ((k >> 3) & 2); }

// the same slots with the thread's part hoisted (T >= 16, a multiple of 16): bit 4 of t + T m is bit4(t) ^ bit4(T m)
template<int N>
GR4B200_HD void fftPark(int t, const Cx (&v)[16], Cx* park) {
    constexpr int T = FftGeom<N>::kThreads;
    static_assert(T >= 16, "parking is used for N >= 256");
    if constexpr (T >= 32) {
        Cx* base = park + (t ^ ((t >> 3) & 2));
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            base[T * m] = v[m];
        }
    } else {
        Cx* even = park + t;
        Cx* odd  = park + (t ^ 2);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            (m & 1 ? odd : even)[T * m] = v[m];
        }
    }
}
// slots of bins 4 (g T + t) and 4 (g T + t) + 2, g = 0..3: base + 4 g T
GR4B200_HD int fftParkReadBase(int t, int half) { return (4 * t + 2 * half) ^ ((t >> 1) & 2); }

} // namespace gr4b200
