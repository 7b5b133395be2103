// Register-level FFT building blocks shared by the device kernels (fft.cu) and the host emulation used by the CPU-side
// tests (tests/host_emulation.cu): 4- and 16-point DFTs and the inter-pass twiddle construction.
#pragma once

#include <cuda_runtime.h>

#include <cmath>

#ifndef GR4B200_HD
#define GR4B200_HD __host__ __device__ __forceinline__
#endif

namespace gr4b200 {

constexpr float kSqrtHalf = 0.70710678118654752440f;
constexpr float kCosPi8   = 0.92387953251128675613f; // cos(pi/8)
constexpr float kSinPi8   = 0.38268343236508977173f; // sin(pi/8)

GR4B200_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
GR4B200_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
GR4B200_HD float2 cmul(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }
GR4B200_HD float2 mulMinusJ(float2 a) { return make_float2(a.y, -a.x); } // a * (-j)

// 4-point forward DFT in place: (x0,x1,x2,x3) -> (X0,X1,X2,X3)
GR4B200_HD void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 s02 = cadd(x0, x2), d02 = csub(x0, x2);
    const float2 s13 = cadd(x1, x3), d13 = mulMinusJ(csub(x1, x3));
    x0 = cadd(s02, s13);
    x1 = cadd(d02, d13);
    x2 = csub(s02, s13);
    x3 = csub(d02, d13);
}

// 16-point forward DFT, natural order in, natural order out: n = 4a + b, k = c + 4d
//   y[b][c] = sum_a x[4a+b] W4^(ac);  y[b][c] *= W16^(bc);  X[c+4d] = sum_b y[b][c] W4^(bd)
GR4B200_HD void dft16(float2 (&x)[16]) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        dft4(x[b], x[4 + b], x[8 + b], x[12 + b]); // now x[4c + b] = y[b][c]
    }
    // W16^m = (cos(m pi/8), -sin(m pi/8))
    x[4 * 1 + 1] = cmul(x[4 * 1 + 1], make_float2(kCosPi8, -kSinPi8));                               // m = 1
    x[4 * 1 + 2] = make_float2((x[4 * 1 + 2].x + x[4 * 1 + 2].y) * kSqrtHalf, (x[4 * 1 + 2].y - x[4 * 1 + 2].x) * kSqrtHalf); // m = 2
    x[4 * 1 + 3] = cmul(x[4 * 1 + 3], make_float2(kSinPi8, -kCosPi8));                               // m = 3
    x[4 * 2 + 1] = make_float2((x[4 * 2 + 1].x + x[4 * 2 + 1].y) * kSqrtHalf, (x[4 * 2 + 1].y - x[4 * 2 + 1].x) * kSqrtHalf); // m = 2
    x[4 * 2 + 2] = mulMinusJ(x[4 * 2 + 2]);                                                           // m = 4
    x[4 * 2 + 3] = make_float2((x[4 * 2 + 3].y - x[4 * 2 + 3].x) * kSqrtHalf, -(x[4 * 2 + 3].x + x[4 * 2 + 3].y) * kSqrtHalf); // m = 6
    x[4 * 3 + 1] = cmul(x[4 * 3 + 1], make_float2(kSinPi8, -kCosPi8));                               // m = 3
    x[4 * 3 + 2] = make_float2((x[4 * 3 + 2].y - x[4 * 3 + 2].x) * kSqrtHalf, -(x[4 * 3 + 2].x + x[4 * 3 + 2].y) * kSqrtHalf); // m = 6
    x[4 * 3 + 3] = cmul(x[4 * 3 + 3], make_float2(-kCosPi8, kSinPi8));                               // m = 9
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        dft4(x[4 * c + 0], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]); // now x[4c + d] = X[c + 4d]
    }
    // undo the index transposition: X[k] with k = c + 4d sits at 4c + d
    float2 t;
#define GR4B200_SWAP(i, j) t = x[i]; x[i] = x[j]; x[j] = t;
    GR4B200_SWAP(1, 4)
    GR4B200_SWAP(2, 8)
    GR4B200_SWAP(3, 12)
    GR4B200_SWAP(6, 9)
    GR4B200_SWAP(7, 13)
    GR4B200_SWAP(11, 14)
#undef GR4B200_SWAP
}

// x[k] *= w^k for k = 1..15 given w^1, w^2, w^4, w^8 (table values, each rounded once)
GR4B200_HD void applyPowers(float2 (&x)[16], float2 w1, float2 w2, float2 w4, float2 w8) {
    const float2 w3 = cmul(w2, w1), w5 = cmul(w4, w1), w6 = cmul(w4, w2), w7 = cmul(w4, w3);
    x[1]  = cmul(x[1], w1);
    x[2]  = cmul(x[2], w2);
    x[3]  = cmul(x[3], w3);
    x[4]  = cmul(x[4], w4);
    x[5]  = cmul(x[5], w5);
    x[6]  = cmul(x[6], w6);
    x[7]  = cmul(x[7], w7);
    x[8]  = cmul(x[8], w8);
    x[9]  = cmul(x[9], cmul(w8, w1));
    x[10] = cmul(x[10], cmul(w8, w2));
    x[11] = cmul(x[11], cmul(w8, w3));
    x[12] = cmul(x[12], cmul(w8, w4));
    x[13] = cmul(x[13], cmul(w8, w5));
    x[14] = cmul(x[14], cmul(w8, w6));
    x[15] = cmul(x[15], cmul(w8, w7));
}

// loads that use the streaming / read-only paths on the device and plain loads in the host emulation
GR4B200_HD float2 loadSample(const float2* p) {
#ifdef __CUDA_ARCH__
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
#else
    return *p;
#endif
}
GR4B200_HD float2 loadTable(const float2* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
GR4B200_HD float loadTable(const float* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
GR4B200_HD float4 loadTable(const float4* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}
GR4B200_HD float mulRn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
}

constexpr int kN4096         = 4096;
constexpr int kThreads4096   = 256;
constexpr int kRowStride4096 = 17; // padded row (16 points + 1) for the pass-2 -> pass-3 exchange
constexpr int kN256          = 256;
constexpr int kThreads256    = 256;

// ---- N = 4096, thread t of 256 ------------------------------------------------------------------------------------
// pass 1: n = 256 n1 + t: load (+window), DFT over n1, twiddle W_4096^(k1 t); result x[k1].
// windowT is the window re-laid out per thread: windowT[16 t + n1] = w[256 n1 + t] (four 16-byte loads per thread)
GR4B200_HD void fft4096Pass1(int t, const float2* in, const float* windowT, const float2* powers1, float2 (&x)[16]) {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        x[n1] = loadSample(in + n1 * 256 + t);
    }
    if (windowT != nullptr) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 w = loadTable(reinterpret_cast<const float4*>(windowT + 16 * t) + q);
            x[4 * q + 0]   = make_float2(mulRn(x[4 * q + 0].x, w.x), mulRn(x[4 * q + 0].y, w.x)); // blocks/fourier fft.hpp:155-162
            x[4 * q + 1]   = make_float2(mulRn(x[4 * q + 1].x, w.y), mulRn(x[4 * q + 1].y, w.y));
            x[4 * q + 2]   = make_float2(mulRn(x[4 * q + 2].x, w.z), mulRn(x[4 * q + 2].y, w.z));
            x[4 * q + 3]   = make_float2(mulRn(x[4 * q + 3].x, w.w), mulRn(x[4 * q + 3].y, w.w));
        }
    }
    dft16(x);
    applyPowers(x, loadTable(powers1 + t), loadTable(powers1 + 256 + t), loadTable(powers1 + 512 + t), loadTable(powers1 + 768 + t));
}
GR4B200_HD void fft4096Store1(int t, const float2 (&x)[16], float2* sA) {
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        sA[k1 * 256 + t] = x[k1];
    }
}
// pass 2: thread = (k1, n3) = (t / 16, t % 16): points n2 at sA[k1][16 n2 + n3]; DFT over n2, twiddle W_256^(k2 n3);
// -> sB[row = k1 + 16 k2][n3]
GR4B200_HD void fft4096Pass2(int t, const float2* sA, const float2* powers2, float2* sB) {
    const int k1 = t >> 4, n3 = t & 15;
    float2    x[16];
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) {
        x[n2] = sA[k1 * 256 + n2 * 16 + n3];
    }
    dft16(x);
    applyPowers(x, loadTable(powers2 + n3), loadTable(powers2 + 16 + n3), loadTable(powers2 + 32 + n3), loadTable(powers2 + 48 + n3));
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
        sB[(k1 + 16 * k2) * kRowStride4096 + n3] = x[k2];
    }
}
// pass 3: thread t = k1 + 16 k2 owns row t; x[k3] = X[t + 256 k3]
GR4B200_HD void fft4096Pass3(int t, const float2* sB, float2 (&x)[16]) {
#pragma unroll
    for (int n3 = 0; n3 < 16; ++n3) {
        x[n3] = sB[t * kRowStride4096 + n3];
    }
    dft16(x);
}

// ---- N = 256, lane t of 16 ------------------------------------------------------------------------------------------
GR4B200_HD void fft256Pass1(int t, const float2* in, const float* window, const float2* powers1, float2 (&x)[16]) {
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        x[n1] = loadSample(in + n1 * 16 + t);
    }
    if (window != nullptr) {
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) {
            const float w = loadTable(window + n1 * 16 + t);
            x[n1]         = make_float2(mulRn(x[n1].x, w), mulRn(x[n1].y, w));
        }
    }
    dft16(x);
    applyPowers(x, loadTable(powers1 + t), loadTable(powers1 + 16 + t), loadTable(powers1 + 32 + t), loadTable(powers1 + 48 + t));
}
GR4B200_HD void fft256Store1(int t, const float2 (&x)[16], float2* sRow /* [16][17] */) {
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        sRow[k1 * 17 + t] = x[k1];
    }
}
// lane t = k1 owns row k1; x[k2] = X[k1 + 16 k2]
GR4B200_HD void fft256Pass2(int t, const float2* sRow, float2 (&x)[16]) {
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) {
        x[n2] = sRow[t * 17 + n2];
    }
    dft16(x);
}

// table[j * count + t] = W_n^(2^j t), j = 0..3, computed in double and rounded once
inline void fillPowerTable(float2* table, size_t count, size_t n) {
    for (int j = 0; j < 4; ++j) {
        for (size_t t = 0; t < count; ++t) {
            const double arg     = -2.0 * 3.14159265358979323846 * static_cast<double>(((size_t{1} << j) * t) % n) / static_cast<double>(n);
            table[j * count + t] = make_float2(static_cast<float>(cos(arg)), static_cast<float>(sin(arg)));
        }
    }
}

} // namespace gr4b200
