// Streaming complex<float> FFT (sm_100a): gr::algorithm::FFT<std::complex<float>>::compute
// (algorithm/include/gnuradio-4.0/algorithm/fourier/fft.hpp:113-153, SimdFFT.hpp:491-690) and the FFT block's
// window / magnitude / phase post-processing (blocks/fourier/include/gnuradio-4.0/fourier/fft.hpp:147-250,
// algorithm/.../fourier/fft_common.hpp:22-123) fused around it.
//
// Contract: unnormalised forward DFT X[k] = sum_n x[n] exp(-j 2 pi k n / N), natural order, float arithmetic.
//
// Transforms above 8192 points run as two launches of fftColumnKernel (fft_large.cuh).
// One kernel template covers every power of two N in [16, 8192] (fft_radix.cuh): a transform is owned by T = N/16
// threads, each holding 16 points in packed f32x2 registers in every pass; passes are radix 16, 16, .., N/16^p
// (Stockham autosort), exchanged through padded shared memory. CTAs loop over transforms (grid: see launchRadix) and
// hold 256/T transforms at a time; groups of T threads synchronise on their own named barrier (or __syncwarp for
// T <= 32), so transforms in one CTA do not wait for each other.
//   input : N >= 1024 -> the whole next transform is prefetched by ONE 1-D bulk async copy (cp.async.bulk, "TMA 1-D")
//           into a staging buffer while the current one is computed (mbarrier per group); smaller N -> direct
//           coalesced 8-byte loads in pass 1 (a transform lives in one warp there, warps run ahead of each other).
//   window: multiplied onto the registers of pass 1 (per-thread layout, four 16-byte loads).
//   output: spectrum -> coalesced 8-byte streaming stores from the last pass; block mode -> the spectrum is parked in
//           shared memory in natural order, every thread finishes four CONSECUTIVE bins (magnitude, phase, Re, Im) and
//           writes 16-byte vectors into the four DataSet planes; per-signal min/max reduced per transform on request.
// Inter-pass twiddles W^(e q), q = 1..15, are built from four table entries (W^e, W^2e, W^4e, W^8e; double precision on
// the host, rounded once) with at most three complex products each.
#include <cmath>
#include <cstdlib>
#include <vector>

#include "async_copy.cuh"
#include "common.cuh"
#include "fft_epilogue.cuh"
#include "fft_large.cuh"
#include "fft_plan.cuh"
#include "fft_radix.cuh"

namespace gr4b200 {
namespace {

struct FftArgs {
    const float2* in;      // batch * N complex samples (or nullptr)
    const float*  inReal;  // batch * N real samples (SpectrumOfReal mode only)
    float2*       out;     // batch * N (c2c) or nullptr
    const float*  windowT; // per-thread window layout windowT[16 t + m] = w[t + T m], or nullptr
    const float2* tables;  // FftGeom<N>::kTableEntries twiddles
    float*        signals; // block mode: [batch][4][N] or nullptr
    float*        ranges;  // block mode: [batch][4][2] or nullptr
    long long     batch;
    unsigned      flags;
};

// SpectrumOfReal: real input (imaginary part zero), full N-bin spectrum out. BlockOfReal: the FFT block on a real stream
// (fft.hpp:147-250 with computeFullSpectrum == false): planes of N/2 values, magnitude and phase of bins [0, N/2) without
// the fft-shift, Re and Im of the LAST N/2 bins of the spectrum (createDataset copies `std::span{_outData}.last(N)`).
enum class Output { Spectrum, Block, SpectrumOfReal, BlockOfReal };
__host__ __device__ constexpr bool isBlockOutput(Output m) { return m == Output::Block || m == Output::BlockOfReal; }
__host__ __device__ constexpr bool isRealInput(Output m) { return m == Output::SpectrumOfReal || m == Output::BlockOfReal; }

template<int T, int Cta>
__device__ __forceinline__ void groupSync(int tr) {
    if constexpr (T <= 32) {
        __syncwarp();
    } else if constexpr (T == Cta) {
        __syncthreads();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(tr + 1), "n"(T) : "memory");
    }
}

// shared-memory plan of one CTA (bytes), shared by the kernel and its launcher
template<int N, Output Mode, bool Tma>
struct FftSmem {
    using G                          = FftGeom<N>;
    static constexpr bool kPingPong  = G::kThreads > 32 && !(N == 8192 && !Tma); // 8192 direct: one 70 KB array, two CTAs per SM
    static constexpr bool kExchange  = G::kPasses >= 2 || (isBlockOutput(Mode) && G::kThreads >= 16);
    static constexpr int  kArray     = G::kPerCta * G::kPadded * 8;
    static constexpr int  kStageOff  = 0;
    static constexpr int  kStage     = Tma ? G::kPerCta * N * 8 : 0;
    static constexpr int  kFirstOff  = kStageOff + kStage;
    static constexpr int  kSecondOff = kFirstOff + (kExchange ? kArray : 0);
    static constexpr int  kRedOff    = kSecondOff + (kPingPong ? kArray : 0);
    static constexpr int  kRed       = (G::kCta / 32) * 8 * 4; // per warp: {lo, hi} x 4 signals
    static constexpr int  kBarOff    = kRedOff + kRed;
    static constexpr int  kTotal     = kBarOff + G::kPerCta * 8;
};

// resident CTAs the register allocation is sized for: bulk staging (100 KB of shared memory at N <= 4096) leaves room for
// two CTAs, the direct-load variant (<= 68 KB) for three; N = 8192 (512 threads, 136+ KB) runs one CTA per SM
template<int N, bool Tma>
constexpr int fftMinCtas() {
    return N > 4096 ? (Tma ? 1 : 2) : (Tma ? 2 : 3);
}

template<int N, Output Mode, bool Tma>
__global__ void __launch_bounds__(FftGeom<N>::kCta, fftMinCtas<N, Tma>()) fftRadixKernel(FftArgs args) {
    using G                  = FftGeom<N>;
    using S                  = FftSmem<N, Mode, Tma>;
    constexpr int  T         = G::kThreads;
    constexpr int  Cta       = G::kCta;
    constexpr int  PerCta    = G::kPerCta;
    constexpr int  Passes    = G::kPasses;
    constexpr bool kPingPong = S::kPingPong;
    constexpr bool kPark     = isBlockOutput(Mode) && T >= 16;
    constexpr bool kRealIn   = isRealInput(Mode);
    constexpr bool kHalf     = Mode == Output::BlockOfReal; // planes of N/2 values
    constexpr int  kPlane    = kHalf ? N / 2 : N;
    constexpr int  kInBytes  = kRealIn ? 4 : 8; // bytes per input sample
    static_assert(!Tma || T >= 64, "bulk staging is used for N >= 1024 only");
    static_assert(!kPingPong || Passes >= 3, "multi-warp transforms have at least three passes");

    extern __shared__ __align__(128) unsigned char smemRaw[];
    const int t  = threadIdx.x % T; // thread within the transform
    const int tr = threadIdx.x / T; // transform within the CTA
    Cx*       stage  = reinterpret_cast<Cx*>(smemRaw + S::kStageOff) + tr * N;
    Cx*       first  = reinterpret_cast<Cx*>(smemRaw + S::kFirstOff) + tr * G::kPadded;
    Cx*       second = reinterpret_cast<Cx*>(smemRaw + S::kSecondOff) + tr * G::kPadded;
    float*    sRed   = reinterpret_cast<float*>(smemRaw + S::kRedOff);
    uint64_t* bar    = reinterpret_cast<uint64_t*>(smemRaw + S::kBarOff) + tr;

    const long long groupStride = static_cast<long long>(gridDim.x) * PerCta;
    const float2* __restrict__ tables = args.tables;
    const bool dB  = (args.flags & GR4B200_FFT_OUTPUT_IN_DB) != 0;
    const bool deg = (args.flags & GR4B200_FFT_OUTPUT_IN_DEG) != 0;

    gridDependencyLaunch(); // common.cuh: the next kernel of the stream may be set up while this one runs
    if constexpr (Tma) {
        if (t == 0) {
            mbarInit(bar, 1);
            fenceBarrierInit();
        }
        groupSync<T, Cta>(tr);
    }
    gridDependencyWait(); // the samples (and the space the planes go to) belong to the previous kernel until here
    if constexpr (Tma) {
        const long long firstXf = static_cast<long long>(blockIdx.x) * PerCta + tr;
        if (t == 0 && firstXf < args.batch) {
            mbarExpectTx(bar, N * kInBytes);
            bulkLoad(stage, kRealIn ? static_cast<const void*>(args.inReal + firstXf * N) : static_cast<const void*>(args.in + firstXf * N), N * kInBytes, bar);
        }
    }
    uint32_t parity = 0;

    // everything a thread needs from the tables depends on t only: with the staged variant (two resident CTAs, 128
    // registers per thread) it is loaded once; the direct-load variant (three CTAs, 80 registers) reloads it from L1
    constexpr bool kHoist    = Tma;
    const bool     hasWindow = args.windowT != nullptr;
    Cx             tw1[kFftTwiddleRegs], tw2[kFftTwiddleRegs], tw3[kFftTwiddleRegs];
    float          win[16];
    if constexpr (kHoist) {
        if constexpr (Passes >= 2) {
            fftLoadTwiddles<N, 1>(t, tables, tw1);
        }
        if constexpr (Passes >= 3) {
            fftLoadTwiddles<N, 2>(t, tables, tw2);
        }
        if constexpr (Passes >= 4) {
            fftLoadTwiddles<N, 3>(t, tables, tw3);
        }
        if (hasWindow) {
            fftLoadWindow(t, args.windowT, win);
        }
    }

    for (long long base = static_cast<long long>(blockIdx.x) * PerCta; base < args.batch; base += groupStride) {
        const long long xf     = base + tr;
        const bool      active = xf < args.batch;
        if constexpr (T >= 32) {
            if (!active) {
                continue; // a whole group (>= one warp) with its own barrier: nothing else waits for it
            }
        }
        Cx v[16];
        // ---- pass 0: load (+ window), radix 16 ---------------------------------------------------------------------------
        if constexpr (Tma) {
            mbarWait(bar, parity);
            parity ^= 1u;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                if constexpr (kRealIn) {
                    v[m] = cxMake(reinterpret_cast<const float*>(stage)[t + T * m], 0.f);
                } else {
                    v[m] = stage[t + T * m];
                }
            }
        } else if constexpr (kRealIn) {
            const float* __restrict__ in = args.inReal + xf * N;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                v[m] = cxMake((T >= 32 || active) ? __ldg(in + t + T * m) : 0.f, 0.f);
            }
        } else {
            const float2* __restrict__ in = args.in + xf * N;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                float2 s = make_float2(0.f, 0.f);
                if (T >= 32 || active) {
                    s = T >= 16 ? ldStream2(in + t + T * m) : __ldg(in + t + T * m); // narrow rows: let L1 merge the sectors
                }
                v[m] = cxMake(s.x, s.y);
            }
        }
        if (hasWindow) {
            if constexpr (kHoist) {
                fftApplyWindow(win, v);
            } else {
                fftApplyWindow(t, args.windowT, v);
            }
        }
        fftPassCompute<N, 0>(t, v, tables);

        if constexpr (Passes >= 2) {
            if constexpr (!kPingPong) {
                groupSync<T, Cta>(tr); // readers of the previous transform are done with the array
            }
            fftScatter<N, 0>(t, v, first);
            groupSync<T, Cta>(tr);
            if constexpr (Tma) {
                const long long next = xf + groupStride;
                if (t == 0 && next < args.batch) { // the staging buffer has been consumed by everyone: refill it
                    mbarExpectTx(bar, N * kInBytes);
                    bulkLoad(stage, kRealIn ? static_cast<const void*>(args.inReal + next * N) : static_cast<const void*>(args.in + next * N), N * kInBytes, bar);
                }
            }
            fftGather<N>(t, first, v);
            if constexpr (kHoist) {
                fftPassWithTwiddles<N, 1>(v, tw1);
            } else {
                fftPassCompute<N, 1>(t, v, tables);
            }
        }
        if constexpr (Passes >= 3) {
            if constexpr (kPingPong) {
                fftScatter<N, 1>(t, v, second);
                groupSync<T, Cta>(tr);
                fftGather<N>(t, second, v);
            } else {
                groupSync<T, Cta>(tr);
                fftScatter<N, 1>(t, v, first);
                groupSync<T, Cta>(tr);
                fftGather<N>(t, first, v);
            }
            if constexpr (kHoist) {
                fftPassWithTwiddles<N, 2>(v, tw2);
            } else {
                fftPassCompute<N, 2>(t, v, tables);
            }
        }
        if constexpr (Passes >= 4) {
            if constexpr (!kPingPong) {
                groupSync<T, Cta>(tr); // single array: pass 2 has been gathered from it by everybody
            }
            fftScatter<N, 2>(t, v, first);
            groupSync<T, Cta>(tr);
            fftGather<N>(t, first, v);
            if constexpr (kHoist) {
                fftPassWithTwiddles<N, 3>(v, tw3);
            } else {
                fftPassCompute<N, 3>(t, v, tables);
            }
        }
        // now v[m] = X[t + T m]; the array read last is `second` for 3 passes, `first` otherwise
        if constexpr (kRealIn) { // the spectrum of a real signal: DC and Nyquist are real (fft.hpp:245-249 sets them so)
            if (t == 0) {
                v[0] = cxMake(cxRe(v[0]), 0.f); // bin 0
                v[8] = cxMake(cxRe(v[8]), 0.f); // bin t + T * 8 = N / 2
            }
        }
        if constexpr (!isBlockOutput(Mode)) {
            float2* __restrict__ out = args.out + xf * N;
            if (T >= 32 || active) {
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    float re, im;
                    cxSplit(v[m], re, im);
                    stStream2(out + t + T * m, make_float2(re, im));
                }
            }
            if constexpr (kPingPong && Passes == 4) { // next pass 0 must not write the array still being read
                Cx* tmp = first;
                first   = second;
                second  = tmp;
            }
        } else {
            float* __restrict__ sig = args.signals + xf * 4 * kPlane;
            float lo[4] = {INFINITY, INFINITY, INFINITY, INFINITY}, hi[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
            const bool wantRanges = args.ranges != nullptr;
            if constexpr (kPark) {
                // park the spectrum in natural order in the array that is idle now
                Cx* park = (kPingPong && Passes == 3) ? first : (kPingPong ? second : first);
                if constexpr (!kPingPong) {
                    groupSync<T, Cta>(tr); // single array: everybody has gathered from it
                }
                fftPark<N>(t, v, park);
                groupSync<T, Cta>(tr);
                if constexpr (kHalf) {
                    fftBlockEpilogueReal<N>(t, park, sig, dB, deg, wantRanges, lo, hi, T >= 32 || active);
                } else {
                    fftBlockEpilogue<N>(t, park, sig, dB, deg, wantRanges, lo, hi, T >= 32 || active);
                }
                if constexpr (kPingPong && Passes == 3) { // parked in `first`: the next pass 0 goes to the other array
                    Cx* tmp = first;
                    first   = second;
                    second  = tmp;
                }
            } else if constexpr (kHalf) { // narrow transforms, real input: registers m < 8 are bins [0, N/2), m >= 8 the upper half
#pragma unroll
                for (int m = 0; m < 8; m += 2) {
                    float mag[2], ph[2];
                    magnitudePhase2(v[m], v[m + 1], 2.f / N, mag[0], mag[1], ph[0], ph[1]);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float re, im;
                        cxSplit(v[m + e + 8], re, im);
                        mag[e] = dB ? decibel(mag[e]) : mag[e];
                        ph[e]  = deg ? toDegrees(ph[e]) : ph[e];
                        lo[0] = fminf(lo[0], mag[e]), hi[0] = fmaxf(hi[0], mag[e]);
                        lo[1] = fminf(lo[1], ph[e]), hi[1] = fmaxf(hi[1], ph[e]);
                        lo[2] = fminf(lo[2], re), hi[2] = fmaxf(hi[2], re);
                        lo[3] = fminf(lo[3], im), hi[3] = fmaxf(hi[3], im);
                        const int k = t + T * (m + e);
                        if (active) {
                            sig[k]              = mag[e];
                            sig[kPlane + k]     = ph[e];
                            sig[2 * kPlane + k] = re; // X[N/2 + k]
                            sig[3 * kPlane + k] = im;
                        }
                    }
                }
            } else {
#pragma unroll
                for (int m = 0; m < 16; m += 2) {
                    float mag[2], ph[2], re[2], im[2];
                    magnitudePhase2(v[m], v[m + 1], 2.f / N, mag[0], mag[1], ph[0], ph[1]);
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        cxSplit(v[m + e], re[e], im[e]);
                        mag[e] = dB ? decibel(mag[e]) : mag[e];
                        ph[e]  = deg ? toDegrees(ph[e]) : ph[e];
                        lo[0] = fminf(lo[0], mag[e]), hi[0] = fmaxf(hi[0], mag[e]);
                        lo[1] = fminf(lo[1], ph[e]), hi[1] = fmaxf(hi[1], ph[e]);
                        lo[2] = fminf(lo[2], re[e]), hi[2] = fmaxf(hi[2], re[e]);
                        lo[3] = fminf(lo[3], im[e]), hi[3] = fmaxf(hi[3], im[e]);
                        const int k       = t + T * (m + e);
                        const int shifted = (k + N / 2) & (N - 1);
                        if (active) {
                            sig[shifted]         = mag[e];
                            sig[N + shifted]     = ph[e];
                            sig[2 * N + k]       = re[e];
                            sig[3 * N + k]       = im[e];
                        }
                    }
                }
            }
            if (wantRanges) { // uniform per launch
                // reduce {lo, hi} x 4 over the T threads of the transform
#pragma unroll
                for (int s = 0; s < 4; ++s) {
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        if (off < T) {
                            lo[s] = fminf(lo[s], __shfl_xor_sync(0xffffffffu, lo[s], off));
                            hi[s] = fmaxf(hi[s], __shfl_xor_sync(0xffffffffu, hi[s], off));
                        }
                    }
                }
                if constexpr (T > 32) {
                    constexpr int Warps = T / 32;
                    float*        mine  = sRed + (tr * Warps) * 8;
                    if ((t & 31) == 0) {
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            mine[(t / 32) * 8 + 2 * s]     = lo[s];
                            mine[(t / 32) * 8 + 2 * s + 1] = hi[s];
                        }
                    }
                    groupSync<T, Cta>(tr);
                    if (t < 4) {
                        float l = INFINITY, h = -INFINITY;
                        for (int w = 0; w < Warps; ++w) {
                            l = fminf(l, mine[w * 8 + 2 * t]);
                            h = fmaxf(h, mine[w * 8 + 2 * t + 1]);
                        }
                        args.ranges[(xf * 4 + t) * 2]     = l;
                        args.ranges[(xf * 4 + t) * 2 + 1] = h;
                    }
                    groupSync<T, Cta>(tr); // sRed is rewritten by the next transform
                } else if (t == 0 && active) {
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        args.ranges[(xf * 4 + s) * 2]     = lo[s];
                        args.ranges[(xf * 4 + s) * 2 + 1] = hi[s];
                    }
                }
            }
        }
    }
}

// per-signal {min, max} recomputed from the planes (after phase unwrapping): one warp per (transform, signal)
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

// fft_common.hpp:72-90: the reference unwraps the natural-order radian phase BEFORE the degree conversion and the
// fft-shift (fft_common.hpp:109-121). One thread per transform (sequential by definition) recomputes the phase from the
// Re / Im planes with the library atan2, unwraps, converts and writes the shifted plane.
__global__ void unwrapPhaseKernel(float* __restrict__ signals, long long batch, int n, int deg) {
    const long long xf = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (xf >= batch) {
        return;
    }
    float*       phase = signals + xf * 4 * n + n;
    const float* re    = signals + xf * 4 * n + 2 * n;
    const float* im    = re + n;
    const float  pi    = 3.14159265358979323846f;
    const int    up    = n - n / 2; // fft-shift = rotate left by n/2: bin k lands at (k + n - n/2) mod n (any n, not only 2^k)
    float        prev  = atan2f(im[0], re[0]);
    phase[up < n ? up : 0] = deg ? toDegrees(prev) : prev;
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
        prev          = cur;
        const int pos = k + up;
        phase[pos >= n ? pos - n : pos] = deg ? toDegrees(cur) : cur;
    }
}

// real-input block with unwrapPhase: the phase plane holds bins [0, N/2) in natural order (radians): unwrap it in place
// (fft_common.hpp:72-90), then convert to degrees if asked; one thread per transform
__global__ void unwrapHalfPlaneKernel(float* __restrict__ signals, long long batch, int half, int deg) {
    const long long xf = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (xf >= batch) {
        return;
    }
    float*      phase = signals + xf * 4 * half + half;
    const float pi    = 3.14159265358979323846f;
    float       prev  = phase[0];
    phase[0]          = deg ? toDegrees(prev) : prev;
    for (int k = 1; k < half; ++k) {
        float cur  = phase[k];
        float diff = __fsub_rn(cur, prev);
        while (diff > pi) {
            cur  = __fsub_rn(cur, __fmul_rn(2.f, pi));
            diff = __fsub_rn(cur, prev);
        }
        while (diff < -pi) {
            cur  = __fadd_rn(cur, __fmul_rn(2.f, pi));
            diff = __fsub_rn(cur, prev);
        }
        prev     = cur;
        phase[k] = deg ? toDegrees(cur) : cur;
    }
}

// ---- n > 8192: two passes of column transforms (fft_large.cuh) -------------------------------------------------------
// Planes: the second step writes the FFT block's four planes straight from registers (lane = column = consecutive bin:
// 64-byte row segments per plane) instead of the spectrum.
#ifndef GR4B200_FFT_COLUMN_THREADS_PER_SM
#define GR4B200_FFT_COLUMN_THREADS_PER_SM 1024 // 256- and 512-point columns at 64 registers: 32 resident warps instead of 16-24 (+3 % / +16 %, r01z_time_fft_large*.jsonl)
#endif
template<int L, bool First, bool Planes>
__global__ void __launch_bounds__(FftColumnGeom<L>::kThreads, (L == 128 ? 4 : GR4B200_FFT_COLUMN_THREADS_PER_SM / L)) fftColumnKernel(FftColumnArgs a, float* __restrict__ signals, unsigned flags) {
    using G = FftColumnGeom<L>;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    Cx*             smem = reinterpret_cast<Cx*>(smemRaw);
    const int       tid  = threadIdx.x;
    const long long tile = blockIdx.x;
    Cx              v[16];
    fftColumnPhaseLoad<L, First>(tid, tile, a, smem, v);
    __syncthreads();
    fftColumnPhasePass<L, 1>(tid, a, smem, v);
    if constexpr (G::kPasses == 3) {
        __syncthreads();
        fftColumnPhaseScatter<L, 1>(tid, v, smem);
        __syncthreads();
        fftColumnPhasePass<L, 2>(tid, a, smem, v);
    }
    if constexpr (First) {
        __syncthreads(); // everybody has gathered: the regions are free for the transposing store
        fftColumnTwiddlePark<L>(tid, tile, a, v, smem);
        __syncthreads();
        fftColumnStoreTransposed<L>(tid, tile, a, smem);
    } else if constexpr (!Planes) {
        fftColumnStoreRows<L>(tid, tile, a, v);
    } else {
        const int       tr = tid & 15, t = tid >> 4;
        const int       tiles = a.cols / 16;
        const long long big   = tile / tiles;
        const int       c     = static_cast<int>(tile % tiles) * 16 + tr;
        const long long n     = static_cast<long long>(L) * a.cols;
        float*          sig   = signals + big * 4 * n;
        const bool      dB = (flags & GR4B200_FFT_OUTPUT_IN_DB) != 0, deg = (flags & GR4B200_FFT_OUTPUT_IN_DEG) != 0;
        const float     scale = 2.f / static_cast<float>(n);
        if (a.realSpectrum != 0) {
            // real input (fft.hpp:147-250 with computeFullSpectrum == false): planes of n/2 values -- magnitude and phase of
            // bins [0, n/2) in natural order (registers m < 8: rows below L/2), Re / Im of bins [n/2, n) (m >= 8)
            const long long half = n / 2;
            sig                  = signals + big * 4 * half;
            if (c == 0 && t == 0) { // bins 0 and n/2 of a real signal's spectrum are real
                v[0] = cxMake(cxRe(v[0]), 0.f);
                v[8] = cxMake(cxRe(v[8]), 0.f);
            }
#pragma unroll
            for (int m = 0; m < 8; m += 2) {
                float mag[2], ph[2];
                magnitudePhase2(v[m], v[m + 1], scale, mag[0], mag[1], ph[0], ph[1]);
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    float re, im;
                    cxSplit(v[m + e + 8], re, im);
                    const long long k = static_cast<long long>(t + G::kT * (m + e)) * a.cols + c;
                    sig[k]            = dB ? decibel(mag[e]) : mag[e];
                    sig[half + k]     = deg ? toDegrees(ph[e]) : ph[e];
                    sig[2 * half + k] = re; // X[n/2 + k]
                    sig[3 * half + k] = im;
                }
            }
            return;
        }
#pragma unroll
        for (int m = 0; m < 16; m += 2) {
            float mag[2], ph[2];
            magnitudePhase2(v[m], v[m + 1], scale, mag[0], mag[1], ph[0], ph[1]);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                float re, im;
                cxSplit(v[m + e], re, im);
                const long long k       = static_cast<long long>(t + G::kT * (m + e)) * a.cols + c;
                const long long shifted = k ^ (n / 2);
                sig[shifted]     = dB ? decibel(mag[e]) : mag[e];
                sig[n + shifted] = deg ? toDegrees(ph[e]) : ph[e];
                sig[2 * n + k]   = re;
                sig[3 * n + k]   = im;
            }
        }
    }
}

} // namespace
} // namespace gr4b200

#include "fft_bluestein.cuh"

using namespace gr4b200;

namespace {

template<int N, Output Mode, bool Tma>
int launchRadix(cudaStream_t stream, const FftArgs& args) {
    using G              = FftGeom<N>;
    using S              = FftSmem<N, Mode, Tma>;
    auto         kernel  = fftRadixKernel<N, Mode, Tma>;
    const size_t smem    = S::kTotal;
    static int   ctasPerSm[64] = {}; // per device, filled on first use
    int          device  = 0;
    GR4B200_CUDA_TRY(cudaGetDevice(&device));
    if (device < 0 || device >= 64) {
        return fail("fft: device index out of range");
    }
    if (ctasPerSm[device] == 0) {
        GR4B200_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        int resident = 0;
        GR4B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kernel, G::kCta, smem));
        ctasPerSm[device] = resident < 1 ? 1 : resident;
    }
    const long long groups = ceilDiv<long long>(args.batch, G::kPerCta);
    // Grid size in units of the resident grid (SMs x CTAs per SM); 0 = one CTA per group of transforms. CTAs that stride
    // through memory in lockstep for the whole launch lose 2-10 % of the bandwidth to the hardware CTA scheduler handing
    // out shorter-lived CTAs (profiles/r01y_time_grid_variants.jsonl): the direct-load sizes take one CTA per group
    // (N = 256: 434 GS/s against 383), the staged sizes keep a loop long enough to amortise their prefetch prologue
    // (N = 4096: x4, block mode 268 GS/s against 262; N = 1024: x16, 421 against 398). GR4B200_FFT_GRID_MULT overrides.
    // (N <= 64: narrow rows, LSU bound, resident grid; N = 8192: one 200 KB CTA per SM, resident grid; its direct-load
    // fallback keeps ONE exchange array, two 70 KB CTAs per SM at 64 registers, one CTA per transform: 309 GS/s against
    // 229 with the ping-pong pair and one CTA per SM, profiles/r01z_time_fft8192_variants.jsonl)
    constexpr int    kDefaultMult = !Tma ? ((N >= 128 && N <= 512) || N == 8192 ? 0 : 1) : (N <= 2048 ? 16 : (N == 4096 ? 4 : 1));
    static const int gridMult     = [] { const char* e = std::getenv("GR4B200_FFT_GRID_MULT"); return e != nullptr ? std::atoi(e) : kDefaultMult; }();
    const long long  cap          = gridMult > 0 ? static_cast<long long>(smCount()) * ctasPerSm[device] * gridMult : groups;
    const int       grid   = static_cast<int>(groups < cap ? groups : cap);
    const cudaError_t launched = launchDependent(kernel, dim3(static_cast<unsigned>(grid)), dim3(G::kCta), smem, stream, args);
    if (launched != cudaSuccess) {
        return checkCuda(launched, "fftRadixKernel");
    }
    return checkLaunch("fftRadixKernel");
}

template<int N, Output Mode>
int launchSize(const gr4b200_fft_plan* plan, cudaStream_t stream, const FftArgs& args) {
    if constexpr (N >= 1024) {
        // bulk copies need 16-byte aligned sources; transforms are N * 8 bytes apart, so the base decides
        const void* source = isRealInput(Mode) ? static_cast<const void*>(args.inReal) : static_cast<const void*>(args.in);
        if (plan->useTma && reinterpret_cast<uintptr_t>(source) % 16 == 0) {
            return launchRadix<N, Mode, true>(stream, args);
        }
    }
    return launchRadix<N, Mode, false>(stream, args);
}

template<Output Mode>
int launchFft(const gr4b200_fft_plan* plan, cudaStream_t stream, FftArgs args) {
    args.windowT = plan->windowT;
    args.tables  = plan->tables;
    switch (plan->n) {
    case 16: return launchSize<16, Mode>(plan, stream, args);
    case 32: return launchSize<32, Mode>(plan, stream, args);
    case 64: return launchSize<64, Mode>(plan, stream, args);
    case 128: return launchSize<128, Mode>(plan, stream, args);
    case 256: return launchSize<256, Mode>(plan, stream, args);
    case 512: return launchSize<512, Mode>(plan, stream, args);
    case 1024: return launchSize<1024, Mode>(plan, stream, args);
    case 2048: return launchSize<2048, Mode>(plan, stream, args);
    case 4096: return launchSize<4096, Mode>(plan, stream, args);
    case 8192: return launchSize<8192, Mode>(plan, stream, args);
    default: return fail("fft: unsupported size");
    }
}


template<int L, bool First, bool Planes>
int launchColumns(cudaStream_t stream, const FftColumnArgs& a, float* signals, unsigned flags) {
    using G            = FftColumnGeom<L>;
    auto       kernel  = fftColumnKernel<L, First, Planes>;
    static bool configured[64] = {};
    int         device = 0;
    GR4B200_CUDA_TRY(cudaGetDevice(&device));
    if (device < 0 || device >= 64) {
        return fail("fft: device index out of range");
    }
    if (!configured[device]) {
        GR4B200_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G::kSmem));
        configured[device] = true;
    }
    const long long tiles = a.batch * (a.cols / 16);
    if (tiles > 0x7fffffffLL) {
        return fail("fft: too many column tiles in one launch");
    }
    kernel<<<static_cast<unsigned>(tiles), G::kThreads, G::kSmem, stream>>>(a, signals, flags);
    return checkLaunch("fftColumnKernel");
}

template<bool First, bool Planes>
int launchColumnsOf(size_t length, cudaStream_t stream, const FftColumnArgs& a, float* signals, unsigned flags) {
    switch (length) {
    case 128: return launchColumns<128, First, Planes>(stream, a, signals, flags);
    case 256: return launchColumns<256, First, Planes>(stream, a, signals, flags);
    case 512: return launchColumns<512, First, Planes>(stream, a, signals, flags);
    default: return fail("fft: unsupported column length");
    }
}

constexpr size_t kLargeSliceSamples = size_t{1} << 26; // scratch of one slice: 512 MiB

// spectrum (signals == nullptr) or planes of `batch` transforms of plan->n > 8192 points
int launchLargeFft(gr4b200_fft_plan* plan, cudaStream_t stream, const float2* in, const float* inReal, float2* out, float* signals, unsigned flags, size_t batch) {
    const size_t n     = plan->n;
    const size_t slice = kLargeSliceSamples / n; // transforms per slice (>= 256)
    const size_t need  = (batch < slice ? batch : slice) * n;
    if (plan->scratchSize < need) {
        GR4B200_CUDA_TRY(cudaStreamSynchronize(stream)); // earlier launches may still read the old scratch
        cudaFree(plan->scratch);
        plan->scratch     = nullptr;
        plan->scratchSize = 0;
        GR4B200_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&plan->scratch), need * sizeof(float2)));
        plan->scratchSize = need;
    }
    for (size_t done = 0; done < batch; done += slice) {
        const size_t  count = batch - done < slice ? batch - done : slice;
        FftColumnArgs first{};
        first.in      = in != nullptr ? reinterpret_cast<const Cx*>(in + done * n) : nullptr;
        first.inReal  = in != nullptr ? nullptr : inReal + done * n;
        first.out     = reinterpret_cast<Cx*>(plan->scratch);
        first.window  = plan->windowN;
        first.twiddle = plan->twiddleN;
        first.tables  = plan->tables1;
        first.cols    = static_cast<int>(plan->n2);
        first.batch   = static_cast<long long>(count);
        int status    = launchColumnsOf<true, false>(plan->n1, stream, first, nullptr, 0);
        if (status != GR4B200_OK) {
            return status;
        }
        FftColumnArgs second{};
        second.in     = reinterpret_cast<const Cx*>(plan->scratch);
        second.out    = signals != nullptr ? nullptr : reinterpret_cast<Cx*>(out + done * n);
        second.tables = plan->tables2;
        second.cols   = static_cast<int>(plan->n1);
        second.batch  = static_cast<long long>(count);
        second.realSpectrum = in == nullptr ? 1 : 0;
        status        = signals != nullptr ? launchColumnsOf<false, true>(plan->n2, stream, second, signals + done * (in == nullptr ? 2 : 4) * n, flags) : launchColumnsOf<false, false>(plan->n2, stream, second, nullptr, 0);
        if (status != GR4B200_OK) {
            return status;
        }
    }
    return GR4B200_OK;
}

template<int N>
void fillTablesFor(std::vector<float2>& table) {
    table.assign(FftGeom<N>::kTableEntries > 0 ? FftGeom<N>::kTableEntries : 1, make_float2(1.f, 0.f));
    fftFillTables<N>(table.data());
}

void fillTables(size_t n, std::vector<float2>& table) {
    switch (n) {
    case 16: return fillTablesFor<16>(table);
    case 32: return fillTablesFor<32>(table);
    case 64: return fillTablesFor<64>(table);
    case 128: return fillTablesFor<128>(table);
    case 256: return fillTablesFor<256>(table);
    case 512: return fillTablesFor<512>(table);
    case 1024: return fillTablesFor<1024>(table);
    case 2048: return fillTablesFor<2048>(table);
    case 4096: return fillTablesFor<4096>(table);
    default: return fillTablesFor<8192>(table);
    }
}

bool upload(const void* host, size_t bytes, void** device) { return cudaMalloc(device, bytes) == cudaSuccess && cudaMemcpy(*device, host, bytes, cudaMemcpyHostToDevice) == cudaSuccess; }

} // namespace

extern "C" {

gr4b200_fft_plan* gr4b200_fft_plan_create(size_t nfft, const float* window_host) {
    const bool radix = nfft >= 16 && (nfft & (nfft - 1)) == 0;
    if (nfft == 0 || (radix && nfft > static_cast<size_t>(kFftLargeMax)) || (!radix && nfft > kBluesteinMaxN)) {
        fail("fft_plan_create: nfft must be a power of two up to 262144 or any size up to 131072");
        return nullptr;
    }
    auto* plan   = new gr4b200_fft_plan;
    plan->device = currentDevice();
    plan->n      = nfft;
    bool ok    = true;
    if (!radix) { // any other size: Bluestein over a power-of-two plan (fft_bluestein.cuh)
        if (!bluesteinPlanCreate(plan, window_host)) {
            checkCuda(cudaGetLastError(), "fft_plan_create");
            gr4b200_fft_plan_destroy(plan);
            return nullptr;
        }
        return plan;
    }
    if (nfft > 8192) { // two passes of column transforms: tables of both lengths, W_n in double, the window as given
        plan->n2 = static_cast<size_t>(fftLargeSecond(static_cast<int>(nfft)));
        plan->n1 = nfft / plan->n2;
        std::vector<float2> table;
        fillTables(plan->n1, table);
        ok = ok && upload(table.data(), table.size() * sizeof(float2), reinterpret_cast<void**>(&plan->tables1));
        fillTables(plan->n2, table);
        ok = ok && upload(table.data(), table.size() * sizeof(float2), reinterpret_cast<void**>(&plan->tables2));
        std::vector<float2> twiddle(nfft);
        for (size_t j = 0; j < nfft; ++j) {
            const double angle = -2.0 * 3.14159265358979323846 * static_cast<double>(j) / static_cast<double>(nfft);
            twiddle[j]         = make_float2(static_cast<float>(std::cos(angle)), static_cast<float>(std::sin(angle)));
        }
        ok = ok && upload(twiddle.data(), nfft * sizeof(float2), reinterpret_cast<void**>(&plan->twiddleN));
        if (window_host != nullptr) {
            ok = ok && upload(window_host, nfft * sizeof(float), reinterpret_cast<void**>(&plan->windowN));
        }
        if (!ok) {
            checkCuda(cudaGetLastError(), "fft_plan_create");
            gr4b200_fft_plan_destroy(plan);
            return nullptr;
        }
        return plan;
    }
    if (window_host != nullptr) {
        const size_t       threads = nfft / 16;
        std::vector<float> transposed(nfft);
        for (size_t t = 0; t < threads; ++t) {
            for (size_t m = 0; m < 16; ++m) {
                transposed[16 * t + m] = window_host[t + threads * m];
            }
        }
        ok = ok && upload(transposed.data(), nfft * sizeof(float), reinterpret_cast<void**>(&plan->windowT));
    }
    std::vector<float2> table;
    fillTables(nfft, table);
    ok = ok && upload(table.data(), table.size() * sizeof(float2), reinterpret_cast<void**>(&plan->tables));
    const char* tmaEnv = std::getenv("GR4B200_FFT_TMA");
    plan->useTma       = !(tmaEnv != nullptr && tmaEnv[0] == '0');
    if (!ok) {
        checkCuda(cudaGetLastError(), "fft_plan_create");
        gr4b200_fft_plan_destroy(plan);
        return nullptr;
    }
    return plan;
}

int gr4b200_fft_plan_destroy(gr4b200_fft_plan* plan) {
    if (plan == nullptr) {
        return GR4B200_OK;
    }
    cudaFree(plan->windowT);
    cudaFree(plan->tables);
    cudaFree(plan->tables1);
    cudaFree(plan->tables2);
    cudaFree(plan->twiddleN);
    cudaFree(plan->windowN);
    cudaFree(plan->scratch);
    cudaFree(plan->chirpConj);
    cudaFree(plan->chirpSpectrum);
    cudaFree(plan->work);
    cudaFree(plan->spectrum);
    gr4b200_fft_plan_destroy(plan->inner);
    delete plan;
    return GR4B200_OK;
}

size_t gr4b200_fft_plan_size(const gr4b200_fft_plan* plan) { return plan == nullptr ? 0 : plan->n; }

int gr4b200_fft_c2c_cf32(gr4b200_fft_plan* plan, void* stream, const float* in, float* out, size_t batch) {
    if (plan == nullptr) {
        return fail("fft_c2c: null plan");
    }
    if (const int status = checkPlanDevice(plan->device, "fft_c2c"); status != GR4B200_OK) {
        return status;
    }
    if (batch == 0) {
        return GR4B200_OK;
    }
    if (in == nullptr || out == nullptr || reinterpret_cast<uintptr_t>(in) % 8 != 0 || reinterpret_cast<uintptr_t>(out) % 8 != 0) {
        return fail("fft_c2c: null or misaligned buffer");
    }
    if (plan->bluesteinM != 0) {
        return bluesteinSpectrum(plan, asStream(stream), reinterpret_cast<const float2*>(in), nullptr, reinterpret_cast<float2*>(out), batch);
    }
    if (plan->n > 8192) {
        return launchLargeFft(plan, asStream(stream), reinterpret_cast<const float2*>(in), nullptr, reinterpret_cast<float2*>(out), nullptr, 0, batch);
    }
    FftArgs args{};
    args.in    = reinterpret_cast<const float2*>(in);
    args.out   = reinterpret_cast<float2*>(out);
    args.batch = static_cast<long long>(batch);
    return launchFft<Output::Spectrum>(plan, asStream(stream), args);
}

int gr4b200_fft_r2c_f32(gr4b200_fft_plan* plan, void* stream, const float* in, float* out, size_t batch) {
    if (plan == nullptr) {
        return fail("fft_r2c: null plan");
    }
    if (const int status = checkPlanDevice(plan->device, "fft_r2c"); status != GR4B200_OK) {
        return status;
    }
    if (batch == 0) {
        return GR4B200_OK;
    }
    if (in == nullptr || out == nullptr || reinterpret_cast<uintptr_t>(in) % 4 != 0 || reinterpret_cast<uintptr_t>(out) % 8 != 0) {
        return fail("fft_r2c: null or misaligned buffer");
    }
    if (plan->bluesteinM != 0) {
        return bluesteinSpectrum(plan, asStream(stream), nullptr, in, reinterpret_cast<float2*>(out), batch);
    }
    if (plan->n > 8192) { // the column passes with a real first load; the full spectrum comes out
        return launchLargeFft(plan, asStream(stream), nullptr, in, reinterpret_cast<float2*>(out), nullptr, 0, batch);
    }
    FftArgs args{};
    args.inReal = in;
    args.out    = reinterpret_cast<float2*>(out);
    args.batch  = static_cast<long long>(batch);
    return launchFft<Output::SpectrumOfReal>(plan, asStream(stream), args);
}

int gr4b200_fft_block_f32(gr4b200_fft_plan* plan, void* stream, const float* in, size_t batch, unsigned flags, float* signals, float* ranges) {
    if (plan == nullptr) {
        return fail("fft_block_f32: null plan");
    }
    if (const int status = checkPlanDevice(plan->device, "fft_block_f32"); status != GR4B200_OK) {
        return status;
    }
    if (batch == 0) {
        return GR4B200_OK;
    }
    if (in == nullptr || signals == nullptr || reinterpret_cast<uintptr_t>(in) % 4 != 0 || reinterpret_cast<uintptr_t>(signals) % 16 != 0) {
        return fail("fft_block_f32: null or misaligned buffer");
    }
    if (plan->bluesteinM != 0) {
        return bluesteinBlock(plan, asStream(stream), nullptr, in, batch, flags, signals, ranges);
    }
    const bool     unwrap       = (flags & GR4B200_FFT_UNWRAP_PHASE) != 0;
    if (plan->n > 8192) { // half-spectrum planes from the second column pass; unwrapping and ranges as passes over the planes
        const int half   = static_cast<int>(plan->n / 2);
        int       status = launchLargeFft(plan, asStream(stream), nullptr, in, nullptr, signals, unwrap ? (flags & ~GR4B200_FFT_OUTPUT_IN_DEG) : flags, batch);
        if (status == GR4B200_OK && unwrap) {
            unwrapHalfPlaneKernel<<<static_cast<int>(ceilDiv<size_t>(batch, 64)), 64, 0, asStream(stream)>>>(signals, static_cast<long long>(batch), half, (flags & GR4B200_FFT_OUTPUT_IN_DEG) != 0 ? 1 : 0);
            status = checkLaunch("unwrapHalfPlaneKernel");
        }
        if (status == GR4B200_OK && ranges != nullptr) {
            const long long rows = static_cast<long long>(batch) * 4;
            rangesKernel<<<static_cast<int>(ceilDiv<long long>(rows * 32, 256)), 256, 0, asStream(stream)>>>(signals, ranges, rows, half);
            status = checkLaunch("rangesKernel");
        }
        return status;
    }
    FftArgs        args{};
    args.inReal                 = in;
    args.signals                = signals;
    args.ranges                 = unwrap ? nullptr : ranges;
    args.batch                  = static_cast<long long>(batch);
    args.flags                  = unwrap ? (flags & ~GR4B200_FFT_OUTPUT_IN_DEG) : flags;
    const int status            = launchFft<Output::BlockOfReal>(plan, asStream(stream), args);
    if (status != GR4B200_OK || !unwrap) {
        return status;
    }
    const int half = static_cast<int>(plan->n / 2);
    unwrapHalfPlaneKernel<<<static_cast<int>(ceilDiv<size_t>(batch, 64)), 64, 0, asStream(stream)>>>(signals, static_cast<long long>(batch), half, (flags & GR4B200_FFT_OUTPUT_IN_DEG) != 0 ? 1 : 0);
    if (ranges != nullptr) {
        const long long rows = static_cast<long long>(batch) * 4;
        rangesKernel<<<static_cast<int>(ceilDiv<long long>(rows * 32, 256)), 256, 0, asStream(stream)>>>(signals, ranges, rows, half);
    }
    return checkLaunch("unwrapHalfPlaneKernel", ranges != nullptr ? 2u : 1u);
}

int gr4b200_fft_block_cf32(gr4b200_fft_plan* plan, void* stream, const float* in, size_t batch, unsigned flags, float* signals, float* ranges) {
    if (plan == nullptr) {
        return fail("fft_block: null plan");
    }
    if (const int status = checkPlanDevice(plan->device, "fft_block"); status != GR4B200_OK) {
        return status;
    }
    if (batch == 0) {
        return GR4B200_OK;
    }
    if (in == nullptr || signals == nullptr || reinterpret_cast<uintptr_t>(in) % 8 != 0 || reinterpret_cast<uintptr_t>(signals) % 16 != 0) {
        return fail("fft_block: null or misaligned buffer");
    }
    if (plan->bluesteinM != 0) {
        return bluesteinBlock(plan, asStream(stream), reinterpret_cast<const float2*>(in), nullptr, batch, flags, signals, ranges);
    }
    const bool     unwrap      = (flags & GR4B200_FFT_UNWRAP_PHASE) != 0;
    const unsigned kernelFlags = unwrap ? (flags & ~GR4B200_FFT_OUTPUT_IN_DEG) : flags;
    float*         kernelRanges = unwrap ? nullptr : ranges; // with unwrapping the phase plane is rewritten afterwards, ranges follow
    if (plan->n > 8192) { // planes from the second column pass; ranges (and unwrapping) as separate passes over the planes
        int status = launchLargeFft(plan, asStream(stream), reinterpret_cast<const float2*>(in), nullptr, nullptr, signals, kernelFlags, batch);
        if (status == GR4B200_OK && unwrap) {
            unwrapPhaseKernel<<<static_cast<int>(ceilDiv<size_t>(batch, 64)), 64, 0, asStream(stream)>>>(signals, static_cast<long long>(batch), static_cast<int>(plan->n), (flags & GR4B200_FFT_OUTPUT_IN_DEG) != 0 ? 1 : 0);
            status = checkLaunch("unwrapPhaseKernel");
        }
        if (status == GR4B200_OK && ranges != nullptr) {
            const long long rows = static_cast<long long>(batch) * 4;
            rangesKernel<<<static_cast<int>(ceilDiv<long long>(rows * 32, 256)), 256, 0, asStream(stream)>>>(signals, ranges, rows, static_cast<int>(plan->n));
            status = checkLaunch("rangesKernel");
        }
        return status;
    }
    FftArgs        args{};
    args.in          = reinterpret_cast<const float2*>(in);
    args.signals     = signals;
    args.ranges      = kernelRanges;
    args.batch       = static_cast<long long>(batch);
    args.flags       = kernelFlags;
    const int status = launchFft<Output::Block>(plan, asStream(stream), args);
    if (status != GR4B200_OK || !unwrap) {
        return status;
    }
    const int n = static_cast<int>(plan->n);
    unwrapPhaseKernel<<<static_cast<int>(ceilDiv<size_t>(batch, 64)), 64, 0, asStream(stream)>>>(signals, static_cast<long long>(batch), n, (flags & GR4B200_FFT_OUTPUT_IN_DEG) != 0 ? 1 : 0);
    if (ranges != nullptr) {
        const long long rows = static_cast<long long>(batch) * 4;
        rangesKernel<<<static_cast<int>(ceilDiv<long long>(rows * 32, 256)), 256, 0, asStream(stream)>>>(signals, ranges, rows, n);
    }
    return checkLaunch("unwrapPhaseKernel", ranges != nullptr ? 2u : 1u);
}

} // extern "C"
