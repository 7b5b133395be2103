// FIR by overlap-save through the 4096-point transform (mode GR4B200_FIR_OVERLAP_SAVE): the tolerance mode that is bound
// by HBM instead of the fp32 pipe. Included by fir.cu.
//
// The reference's fir_filter::processOne (time_domain_filter.hpp:44-47) costs 2 * nTaps separately rounded operations
// per real sample: 508 flop per 16 B for 127 taps on complex<float>, which the CUDA cores cannot deliver at memory speed
// (31.75 flop/B against a ridge of 11.4). y = x * b through Y = X . H costs ~130 flop per sample at any
// filter length up to 2049 taps. One CTA (256 threads) per block of 4096 input samples, of which the first nTaps - 1 are
// the overlap with the previous block:
//   registers v[m] = x[start + t + 256 m]   (the window is prefetched by one bulk copy while the previous block is being
//                                            transformed; the overlap is rounded up to an even count so that every
//                                            window starts 16-byte aligned)
//   three radix-16 passes (fft_radix.cuh)   -> v[m] = X[t + 256 m]
//   v[m] = conj(v[m] * H[t + 256 m] / N)    (H in registers, a function of the thread only: loaded once per CTA)
//   the same three passes again             -> v[m] = conj(y[t + 256 m])  -- the inverse transform by conjugation; after a
//                                              forward transform register m holds bin t + 256 m, which is exactly the
//                                              gather layout of the next transform's first pass: no exchange in between
//   outputs t + 256 m >= nTaps - 1 are stored (hop = 4097 - nTaps valid outputs per block).
// Arithmetic: float transforms, twiddles and H computed in double and rounded once. NOT the reference's rounding: the
// result differs from the exact mode by a few 1e-7 of sum|b| * max|x| (stated and checked in tests/test_gpu_parity.py);
// x[<0] comes from the plan's carried state like in the other modes.
#pragma once

#include "async_copy.cuh"
#include "fft_radix.cuh"

namespace gr4b200 {
namespace {

constexpr int kOlsN       = 4096;
constexpr int kOlsThreads = 256;
constexpr int kOlsMaxTaps = 2049; // hop >= 2048

struct OlsArgs {
    const float2* in;
    const float2* state;    // haloPad samples in front of in[0]
    float2*       out;
    const float2* spectrum; // [16 t + m] = H[t + 256 m] / 4096
    const float2* tables;   // twiddle tables of FftGeom<4096>
    long long     nIn;
    long long     nBlocks;
    int           overlap;  // nTaps - 1
    int           hop;      // 4096 - overlap
    int           haloPad;
    int           useBulk;  // in 16-byte aligned: windows are prefetched by bulk copies
};

__global__ void __launch_bounds__(kOlsThreads, 2) firOverlapSaveKernel(OlsArgs a) {
    using G = FftGeom<kOlsN>;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ uint64_t                            fullBar;
    Cx* const stage  = reinterpret_cast<Cx*>(smemRaw);            // the next block's 4096 samples, bulk-copied while this one is transformed
    Cx* const arrayA = stage + kOlsN;
    Cx* const arrayB = arrayA + G::kPadded;
    const int t      = threadIdx.x;

    Cx h[16], tw1[kFftTwiddleRegs], tw2[kFftTwiddleRegs];
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        h[m] = cxLoadTable(a.spectrum + 16 * t + m);
    }
    fftLoadTwiddles<kOlsN, 1>(t, a.tables, tw1);
    fftLoadTwiddles<kOlsN, 2>(t, a.tables, tw2);
    const Cx unused[kFftTwiddleRegs] = {};

    auto transform = [&](Cx (&v)[16]) { // forward DFT of the 4096 points held as v[m] = data[t + 256 m]; bins come back the same way
        fftPassWithTwiddles<kOlsN, 0>(v, unused);
        fftScatter<kOlsN, 0>(t, v, arrayA);
        __syncthreads();
        fftGather<kOlsN>(t, arrayA, v);
        fftPassWithTwiddles<kOlsN, 1>(v, tw1);
        fftScatter<kOlsN, 1>(t, v, arrayB);
        __syncthreads();
        fftGather<kOlsN>(t, arrayB, v);
        fftPassWithTwiddles<kOlsN, 2>(v, tw2);
    };
    // a window that lies inside the input, 16-byte aligned (a.hop and a.overlap are even, so every window is if the first is)
    auto staged = [&](long long block) {
        const long long start = block * a.hop - a.overlap;
        return a.useBulk != 0 && start >= 0 && start + kOlsN <= a.nIn;
    };
    auto issue = [&](long long block) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the stage was last read through the generic proxy
        mbarExpectTx(&fullBar, kOlsN * sizeof(Cx));
        bulkLoad(stage, a.in + (block * a.hop - a.overlap), kOlsN * sizeof(Cx), &fullBar);
    };

    if (t == 0) {
        mbarInit(&fullBar, 1);
        fenceBarrierInit();
    }
    __syncthreads();
    long long block = blockIdx.x;
    if (t == 0 && block < a.nBlocks && staged(block)) {
        issue(block);
    }
    uint32_t parity = 0;

    for (; block < a.nBlocks; block += gridDim.x) {
        const long long start = block * a.hop - a.overlap; // stream index of the window's first sample
        Cx              v[16];
        if (staged(block)) {
            mbarWait(&fullBar, parity);
            parity ^= 1u;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                v[m] = stage[t + 256 * m];
            }
        } else { // first block (history from the carried state), last block (zeros past the end), misaligned input
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const long long q = start + t + 256 * m;
                float2          x = make_float2(0.f, 0.f);
                if (q < 0) {
                    x = a.state[a.haloPad + q];
                } else if (q < a.nIn) {
                    x = a.in[q];
                }
                v[m] = cxMake(x.x, x.y);
            }
        }
        __syncthreads(); // the stage has been read by everybody: the next block streams in behind the two transforms
        const long long next = block + gridDim.x;
        if (t == 0 && next < a.nBlocks && staged(next)) {
            issue(next);
        }
        transform(v);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            v[m] = cxMul(v[m], h[m]);
            v[m] = pkMul(v[m], cxMake(1.f, -1.f)); // conj
        }
        transform(v); // (its first scatter into arrayA comes after the barrier that followed the last read of arrayA)
        const long long outBase = block * a.hop - a.overlap; // output index of window position 0
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const int       i = t + 256 * m;
            const long long o = outBase + i;
            if (i >= a.overlap && o < a.nIn) {
                float re, im;
                cxSplit(v[m], re, im);
                stStream2(a.out + o, make_float2(re, -im));
            }
        }
        // the next iteration's first scatter into arrayA: every thread has passed the barrier behind the last gather of arrayA
    }
}

// H[k] / 4096 of the taps, in double on the host, in the kernel's per-thread layout
inline void olsSpectrum(const float* taps, size_t nTaps, std::vector<float2>& layout) {
    std::vector<double> re(kOlsN, 0.0), im(kOlsN, 0.0);
    for (size_t k = 0; k < nTaps; ++k) {
        re[k] = static_cast<double>(taps[k]);
    }
    // iterative radix-2 transform in double (plan creation only)
    for (size_t i = 1, j = 0; i < static_cast<size_t>(kOlsN); ++i) {
        size_t bit = kOlsN >> 1;
        for (; j & bit; bit >>= 1) {
            j ^= bit;
        }
        j ^= bit;
        if (i < j) {
            std::swap(re[i], re[j]);
            std::swap(im[i], im[j]);
        }
    }
    for (size_t len = 2; len <= static_cast<size_t>(kOlsN); len <<= 1) {
        const double angle = -2.0 * 3.14159265358979323846 / static_cast<double>(len);
        for (size_t i = 0; i < static_cast<size_t>(kOlsN); i += len) {
            for (size_t k = 0; k < len / 2; ++k) {
                const double wr = std::cos(angle * static_cast<double>(k)), wi = std::sin(angle * static_cast<double>(k));
                const double ur = re[i + k], ui = im[i + k];
                const double vr = re[i + k + len / 2] * wr - im[i + k + len / 2] * wi;
                const double vi = re[i + k + len / 2] * wi + im[i + k + len / 2] * wr;
                re[i + k]           = ur + vr;
                im[i + k]           = ui + vi;
                re[i + k + len / 2] = ur - vr;
                im[i + k + len / 2] = ui - vi;
            }
        }
    }
    layout.resize(kOlsN);
    for (int t = 0; t < kOlsThreads; ++t) {
        for (int m = 0; m < 16; ++m) {
            const int k        = t + 256 * m;
            layout[16 * t + m] = make_float2(static_cast<float>(re[k] / kOlsN), static_cast<float>(im[k] / kOlsN));
        }
    }
}

inline int launchOverlapSave(cudaStream_t stream, OlsArgs a) {
    a.nBlocks = ceilDiv<long long>(a.nIn, a.hop);
    const size_t smem = (static_cast<size_t>(kOlsN) + 2 * static_cast<size_t>(FftGeom<kOlsN>::kPadded)) * sizeof(Cx);
    static bool  configured[64] = {};
    const int    device = currentDevice();
    if (device < 0 || device >= 64) {
        return fail("fir: device index out of range");
    }
    if (!configured[device]) {
        GR4B200_CUDA_TRY(cudaFuncSetAttribute(firOverlapSaveKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        configured[device] = true;
    }
    // one CTA per block of the stream: a streaming kernel is fastest under the hardware CTA scheduler (see mathop.cu)
    const long long cap  = static_cast<long long>(smCount()) * 2 * 32;
    const int       grid = static_cast<int>(a.nBlocks < cap ? a.nBlocks : cap);
    firOverlapSaveKernel<<<grid, kOlsThreads, smem, stream>>>(a);
    return checkLaunch("firOverlapSaveKernel");
}

} // namespace
} // namespace gr4b200
