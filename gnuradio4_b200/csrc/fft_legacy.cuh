// The first-generation FFT kernels (scalar float2 arithmetic, per-size kernels for 4096 and 256, radix-2 Stockham for
// the rest), kept for A/B timing against the packed radix family in fft.cu (GR4B200_FFT_LEGACY=1). Not the default path.
#pragma once

#include <cmath>
#include <vector>

#include "common.cuh"
#include "fft_core.cuh"

namespace gr4b200 {
namespace legacy {

struct FftArgs {
    const float2* in;      // batch * N
    float2*       out;     // batch * N (c2c) or nullptr
    const float*  window;  // N floats or nullptr (natural order)
    const float*  windowT; // 4096 only: per-thread layout windowT[16 t + n1] = w[256 n1 + t], or nullptr
    const float2* powers1; // [4][N/16]: W_N^(2^j t)
    const float2* powers2; // [4][16]:   W_256^(2^j n3)   (4096 only)
    float*        signals; // block mode: [batch][4][N] or nullptr
    float*        ranges;  // block mode: [batch][4][2] or nullptr
    long long     batch;
    unsigned      flags;
};

enum class Output { Spectrum, Block };

// fft_common.hpp:37-44: hypot(re, im) * 2 / N, optional 20 log10 with log(0) -> lowest().
// N is a power of two here, so (m * 2) / N == m * (2 / N) bit for bit; sqrt(fma(re, re, im*im)) is within 1 ulp of
// hypot whenever the sum of squares stays in the normal range, which is tested first (else: hypotf).
// rarely taken paths are kept out of line: the unrolled epilogue must stay small enough for the instruction cache
__device__ __noinline__ float hypotSlow(float x, float y) { return hypotf(x, y); }
__device__ __noinline__ float atan2Slow(float y, float x) { return atan2f(y, x); }
__device__ __noinline__ float decibel(float mag) { return mag > 0.f ? __fmul_rn(20.f, log10f(mag)) : -3.402823466e+38f; }

__device__ __forceinline__ float magnitudeOf(float2 v, float twoOverN, bool dB) {
    const float sumSq = fmaf(v.x, v.x, v.y * v.y);
    const float norm  = (sumSq > 1.0e-30f && sumSq < 1.0e30f) ? __fsqrt_rn(sumSq) : hypotSlow(v.x, v.y);
    const float mag   = __fmul_rn(norm, twoOverN);
    return dB ? decibel(mag) : mag;
}

// fft_common.hpp:107: atan2(im, re). Octant reduction + the degree-17 odd minimax polynomial of Abramowitz & Stegun
// 4.4.49 (|relative error| <= 2e-8 on [0, 1]): absolute error <= 3e-7 rad, i.e. within 2 ulp of pi-sized phases;
// zeros, infinities and NaNs take the library path so that the special-value table of atan2 holds.
__device__ __forceinline__ float phaseOf(float2 v, bool deg) {
    const float ax = fabsf(v.x), ay = fabsf(v.y);
    const float hi = fmaxf(ax, ay), lo = fminf(ax, ay);
    float       phase;
    if (hi > 1.0e-30f && hi < 1.0e30f) {
        const float t  = __fdividef(lo, hi);
        const float t2 = t * t;
        float       p  = 0.0028662257f;
        p              = fmaf(p, t2, -0.0161657367f);
        p              = fmaf(p, t2, 0.0429096138f);
        p              = fmaf(p, t2, -0.0752896400f);
        p              = fmaf(p, t2, 0.1065626393f);
        p              = fmaf(p, t2, -0.1420889944f);
        p              = fmaf(p, t2, 0.1999355085f);
        p              = fmaf(p, t2, -0.3333314528f);
        p              = fmaf(p * t2, t, t);
        p              = ay > ax ? 1.57079632679489661923f - p : p;
        p              = v.x < 0.f ? 3.14159265358979323846f - p : p;
        phase          = copysignf(p, v.y);
    } else {
        phase = atan2Slow(v.y, v.x);
    }
    return deg ? __fmul_rn(__fmul_rn(phase, 180.f), 0.318309886183790671538f) : phase;
}

// min/max of one value per thread over the CTA slice of `threadsPerTransform` threads, written by its first thread
template<int ThreadsPerTransform>
__device__ __forceinline__ void rangeReduce(float lo, float hi, float* sRed, int laneInTransform, int transformInCta, float* dst) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        if (off < ThreadsPerTransform) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
        }
    }
    if constexpr (ThreadsPerTransform > 32) {
        constexpr int Warps = ThreadsPerTransform / 32;
        const int     warp  = laneInTransform / 32;
        __syncthreads();
        if ((laneInTransform & 31) == 0) {
            sRed[(transformInCta * Warps + warp) * 2 + 0] = lo;
            sRed[(transformInCta * Warps + warp) * 2 + 1] = hi;
        }
        __syncthreads();
        if (laneInTransform == 0) {
            for (int w = 1; w < Warps; ++w) {
                lo = fminf(lo, sRed[(transformInCta * Warps + w) * 2 + 0]);
                hi = fmaxf(hi, sRed[(transformInCta * Warps + w) * 2 + 1]);
            }
        }
    }
    if (laneInTransform == 0) {
        dst[0] = lo;
        dst[1] = hi;
    }
}

// ---- N = 4096 ------------------------------------------------------------------------------------------------------

template<Output Mode>
__global__ void __launch_bounds__(kThreads4096, 3) fft4096Kernel(FftArgs args) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float2* sA   = reinterpret_cast<float2*>(smemRaw);          // [16][256]
    float2* sB   = sA + kN4096;                                  // [256][17]
    float*  sRed = reinterpret_cast<float*>(sB + 256 * kRowStride4096);

    const int t = threadIdx.x;
    for (long long xf = blockIdx.x; xf < args.batch; xf += gridDim.x) {
        const float2* __restrict__ in = args.in + xf * kN4096;
        float2 x[16];
        fft4096Pass1(t, in, args.windowT, args.powers1, x);
        __syncthreads(); // previous transform's pass-2 readers are done with sA
        fft4096Store1(t, x, sA);
        __syncthreads();
        fft4096Pass2(t, sA, args.powers2, sB); // previous transform's pass-3 readers of sB passed the barrier above
        __syncthreads();
        fft4096Pass3(t, sB, x);
        if constexpr (Mode == Output::Spectrum) {
            float2* __restrict__ out = args.out + xf * kN4096;
#pragma unroll
            for (int k3 = 0; k3 < 16; ++k3) {
                stStream2(out + k3 * 256 + t, x[k3]);
            }
        } else {
            const bool dB  = (args.flags & GR4B200_FFT_OUTPUT_IN_DB) != 0;
            const bool deg = (args.flags & GR4B200_FFT_OUTPUT_IN_DEG) != 0;
            float* __restrict__ sig = args.signals + xf * 4 * kN4096;
            float lo[4] = {INFINITY, INFINITY, INFINITY, INFINITY}, hi[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            // sA is idle after pass 2: park the spectrum there in natural order so that every thread can finish four
            // CONSECUTIVE bins and write 16-byte vectors to each of the four planes (4x fewer store instructions)
#pragma unroll
            for (int k3 = 0; k3 < 16; ++k3) {
                sA[k3 * 256 + t] = x[k3];
            }
            __syncthreads();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int    k0      = 4 * (g * 256 + t);
                const int    shifted = (k0 + kN4096 / 2) & (kN4096 - 1); // fft-shift keeps groups of four together
                const float4 a       = *reinterpret_cast<const float4*>(sA + k0);
                const float4 b       = *reinterpret_cast<const float4*>(sA + k0 + 2);
                const float2 v[4]    = {make_float2(a.x, a.y), make_float2(a.z, a.w), make_float2(b.x, b.y), make_float2(b.z, b.w)};
                float        mag[4], ph[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    mag[e] = magnitudeOf(v[e], 2.f / kN4096, dB);
                    ph[e]  = phaseOf(v[e], deg);
                    if (args.ranges != nullptr) {
                        lo[0] = fminf(lo[0], mag[e]), hi[0] = fmaxf(hi[0], mag[e]);
                        lo[1] = fminf(lo[1], ph[e]), hi[1] = fmaxf(hi[1], ph[e]);
                        lo[2] = fminf(lo[2], v[e].x), hi[2] = fmaxf(hi[2], v[e].x);
                        lo[3] = fminf(lo[3], v[e].y), hi[3] = fmaxf(hi[3], v[e].y);
                    }
                }
                stStream4(reinterpret_cast<float4*>(sig + shifted), make_float4(mag[0], mag[1], mag[2], mag[3]));
                stStream4(reinterpret_cast<float4*>(sig + kN4096 + shifted), make_float4(ph[0], ph[1], ph[2], ph[3]));
                stStream4(reinterpret_cast<float4*>(sig + 2 * kN4096 + k0), make_float4(v[0].x, v[1].x, v[2].x, v[3].x));
                stStream4(reinterpret_cast<float4*>(sig + 3 * kN4096 + k0), make_float4(v[0].y, v[1].y, v[2].y, v[3].y));
            }
            if (args.ranges != nullptr) {
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    rangeReduce<kThreads4096>(lo[s], hi[s], sRed, t, 0, args.ranges + (xf * 4 + s) * 2);
                }
            }
        }
    }
}

// ---- N = 256: 16 threads per transform, 16 transforms per CTA ----------------------------------------------------------

template<Output Mode>
__global__ void __launch_bounds__(kThreads256) fft256Kernel(FftArgs args) {
    __shared__ float2 sB[16][16 * 17]; // per transform: [row = k1][17]
    const int t  = threadIdx.x & 15;    // lane within the transform
    const int tr = threadIdx.x >> 4;    // transform within the CTA
    const long long groups = (args.batch + 15) / 16;
    for (long long g = blockIdx.x; g < groups; g += gridDim.x) {
        const long long xf     = g * 16 + tr;
        const bool      active = xf < args.batch;
        float2          x[16];
        if (active) {
            fft256Pass1(t, args.in + xf * kN256, args.window, args.powers1, x);
        }
        __syncthreads();
        if (active) {
            fft256Store1(t, x, sB[tr]);
        }
        __syncthreads();
        if (active) {
            fft256Pass2(t, sB[tr], x);
            if constexpr (Mode == Output::Spectrum) {
                float2* __restrict__ out = args.out + xf * kN256;
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) {
                    stStream2(out + k2 * 16 + t, x[k2]);
                }
            } else {
                const bool dB  = (args.flags & GR4B200_FFT_OUTPUT_IN_DB) != 0;
                const bool deg = (args.flags & GR4B200_FFT_OUTPUT_IN_DEG) != 0;
                float* __restrict__ sig = args.signals + xf * 4 * kN256;
#pragma unroll
                for (int k2 = 0; k2 < 16; ++k2) {
                    const int k       = k2 * 16 + t;
                    const int shifted = (k + kN256 / 2) & (kN256 - 1);
                    sig[shifted]             = magnitudeOf(x[k2], 2.f / kN256, dB);
                    sig[kN256 + shifted]     = phaseOf(x[k2], deg);
                    sig[2 * kN256 + k]       = x[k2].x;
                    sig[3 * kN256 + k]       = x[k2].y;
                }
            }
        }
    }
}

// ---- any power of two in [16, 8192]: shared-memory Stockham radix-2, N/2 threads, twiddles from a W_N^k table ---------
template<Output Mode>
__global__ void fftGenericKernel(FftArgs args, int n, int log2n, const float2* __restrict__ twiddle /* W_N^k, k < N/2 */) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float2* bufA = reinterpret_cast<float2*>(smemRaw);
    float2* bufB = bufA + n;
    const int half = n / 2;
    for (long long xf = blockIdx.x; xf < args.batch; xf += gridDim.x) {
        const float2* __restrict__ in = args.in + xf * n;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            float2 v = in[i];
            if (args.window != nullptr) {
                const float w = args.window[i];
                v             = make_float2(__fmul_rn(v.x, w), __fmul_rn(v.y, w));
            }
            bufA[i] = v;
        }
        __syncthreads();
        float2* src = bufA;
        float2* dst = bufB;
        // Stockham autosort, decimation in frequency: length l halves, stride s doubles
        int s = 1;
        for (int l = half; l >= 1; l >>= 1, s <<= 1) {
            for (int i = threadIdx.x; i < half; i += blockDim.x) {
                const int    p  = i / s;          // 0 .. l-1
                const int    q  = i % s;          // 0 .. s-1
                const float2 w  = twiddle[p * s]; // exp(-j 2 pi p / (2 l)) = W_N^(p * s) since 2 l s = N
                const float2 a  = src[q + s * p];
                const float2 b  = src[q + s * (p + l)];
                dst[q + s * (2 * p)]     = cadd(a, b);
                dst[q + s * (2 * p + 1)] = cmul(csub(a, b), w);
            }
            __syncthreads();
            float2* tmp = src;
            src         = dst;
            dst         = tmp;
        }
        if constexpr (Mode == Output::Spectrum) {
            float2* __restrict__ out = args.out + xf * n;
            for (int i = threadIdx.x; i < n; i += blockDim.x) {
                out[i] = src[i];
            }
        } else {
            const bool dB  = (args.flags & GR4B200_FFT_OUTPUT_IN_DB) != 0;
            const bool deg = (args.flags & GR4B200_FFT_OUTPUT_IN_DEG) != 0;
            float* __restrict__ sig = args.signals + xf * 4 * n;
            for (int k = threadIdx.x; k < n; k += blockDim.x) {
                const int shifted = (k + half) & (n - 1);
                sig[shifted]             = magnitudeOf(src[k], 2.f / static_cast<float>(n), dB);
                sig[n + shifted]         = phaseOf(src[k], deg);
                sig[2 * n + k]           = src[k].x;
                sig[3 * n + k]           = src[k].y;
            }
        }
        __syncthreads();
        (void)log2n;
    }
}

// per-signal {min, max} for sizes whose kernel does not fuse it: one warp per (transform, signal)
__global__ void rangesKernel(const float* __restrict__ signals, float* __restrict__ ranges, long long rows, int n) {
    const long long row  = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / 32;
    const int       lane = threadIdx.x & 31;
    if (row >= rows) {
        return;
    }
    float lo = INFINITY, hi = -INFINITY;
    for (int i = lane; i < n; i += 32) {
        const float v = signals[row * n + i];
        lo            = fminf(lo, v);
        hi            = fmaxf(hi, v);
    }
    for (int off = 16; off > 0; off >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, off));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, off));
    }
    if (lane == 0) {
        ranges[row * 2]     = lo;
        ranges[row * 2 + 1] = hi;
    }
}

// fft_common.hpp:72-90 applied to the (already shifted? no: unshifted) phase plane: the reference unwraps BEFORE the
// degree conversion and the fft-shift (fft_common.hpp:109-121). This kernel therefore runs on the natural-order radian
// phase, one thread per transform (sequential by definition), then re-applies degree conversion and the shift.
__global__ void unwrapPhaseKernel(float* __restrict__ signals, long long batch, int n, int deg) {
    const long long xf = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (xf >= batch) {
        return;
    }
    float*      phase = signals + xf * 4 * n + n;              // shifted plane, radians (kernel wrote it with deg = 0)
    const float* re   = signals + xf * 4 * n + 2 * n;
    const float* im   = re + n;
    const float pi    = 3.14159265358979323846f;
    const int   half  = n / 2;
    float       prev  = atan2f(im[0], re[0]);
    phase[half]       = deg ? __fmul_rn(__fmul_rn(prev, 180.f), 0.318309886183790671538f) : prev;
    for (int k = 1; k < n; ++k) {
        float cur  = atan2f(im[k], re[k]);
        float diff = __fsub_rn(cur, prev);
        while (diff > pi) {
            cur  = __fsub_rn(cur, __fmul_rn(2.f, pi));
            diff = __fsub_rn(cur, prev);
        }
        while (diff < -pi) {
            cur  = __fadd_rn(cur, __fmul_rn(2.f, pi));
            diff = __fsub_rn(cur, prev);
        }
        prev                        = cur;
        phase[(k + half) & (n - 1)] = deg ? __fmul_rn(__fmul_rn(cur, 180.f), 0.318309886183790671538f) : cur;
    }
}

struct LegacyTables {
    size_t  n       = 0;
    int     log2n   = 0;
    float*  window  = nullptr;
    float*  windowT = nullptr;
    float2* powers1 = nullptr;
    float2* powers2 = nullptr;
    float2* twiddle = nullptr;
};

template<Output Mode>
int launchFft(const LegacyTables& plan, cudaStream_t stream, FftArgs args) {
    args.window  = plan.window;
    args.windowT = plan.windowT;
    args.powers1 = plan.powers1;
    args.powers2 = plan.powers2;
    const long long sms = smCount();
    if (plan.n == 4096) {
        const size_t smem = (kN4096 + 256 * kRowStride4096) * sizeof(float2) + 64 * sizeof(float);
        auto         kernel = fft4096Kernel<Mode>;
        GR4B200_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        int ctasPerSm = 0;
        GR4B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, kernel, kThreads4096, smem));
        const long long cap  = sms * (ctasPerSm < 1 ? 1 : ctasPerSm);
        const int       grid = static_cast<int>(args.batch < cap ? args.batch : cap);
        kernel<<<grid, kThreads4096, smem, stream>>>(args);
        return checkLaunch("fft4096Kernel");
    }
    if (plan.n == 256) {
        const long long groups = (args.batch + 15) / 16;
        const long long cap    = sms * 4;
        const int       grid   = static_cast<int>(groups < cap ? groups : cap);
        fft256Kernel<Mode><<<grid, kThreads256, 0, stream>>>(args);
        if (Mode == Output::Block && args.ranges != nullptr) {
            const long long rows = args.batch * 4;
            rangesKernel<<<static_cast<int>(ceilDiv<long long>(rows * 32, 256)), 256, 0, stream>>>(args.signals, args.ranges, rows, 256);
        }
        return checkLaunch("fft256Kernel");
    }
    const int    n       = static_cast<int>(plan.n);
    const int    threads = n / 2 < 32 ? 32 : (n / 2 > 512 ? 512 : n / 2);
    const size_t smem    = 2 * plan.n * sizeof(float2);
    auto         kernel  = fftGenericKernel<Mode>;
    if (smem > 48 * 1024) {
        GR4B200_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    }
    const long long cap  = sms * 8;
    const int       grid = static_cast<int>(args.batch < cap ? args.batch : cap);
    kernel<<<grid, threads, smem, stream>>>(args, n, plan.log2n, plan.twiddle);
    if (Mode == Output::Block && args.ranges != nullptr) {
        const long long rows = args.batch * 4;
        rangesKernel<<<static_cast<int>(ceilDiv<long long>(rows * 32, 256)), 256, 0, stream>>>(args.signals, args.ranges, rows, n);
    }
    return checkLaunch("fftGenericKernel");
}


} // namespace legacy
} // namespace gr4b200
