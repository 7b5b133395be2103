// FFT sizes that are not a power of two (or below 16): Bluestein's chirp-z form on top of the power-of-two kernels.
// Included by fft.cu (it uses that file's plane post-passes and the plan).
//
// Reference: gr::algorithm::FFT::compute picks SimdFFT for sizes that factor into {2, 3, 4, 5} (multiples of 16) and
// its own Bluestein otherwise (algorithm/include/gnuradio-4.0/algorithm/fourier/fft.hpp:113-153, 353-381, 408-425); the
// result it specifies is the same unnormalised forward DFT X[k] = sum_n x[n] e^{-j 2 pi k n / N} in natural order for
// every size. That is what this path returns, for any N in [1, 131072] that is not served by the radix kernels:
//     w[n] = e^{+j pi n^2 / N}:   X[k] = conj(w[k]) * sum_n (x[n] conj(w[n])) w[k - n]
// i.e. a cyclic convolution of length M = 2^ceil(log2(2N - 1)) (>= 16) of a[n] = x[n] conj(w[n]) with the symmetric
// chirp b[n] = b[M - n] = w[n], done with three M-point forward transforms (one of them, FFT(b) / M, precomputed in
// double on the host): c = IFFT(FFT(a) . FFT(b)) = conj(FFT(conj(FFT(a) . FFT(b)))) / M.
// n^2 mod 2N is formed in integers and the chirp in double before rounding (the reference does the same in the element
// type, fft.hpp:410-414). Launches per call: chirp-in, FFT, pointwise, FFT, chirp-out (+ the plane kernel in block mode).
// (The compiled reference's own Bluestein branch returns values that are off by O(1) from the DFT for every size it
// serves -- tests/test_gpu_golden.py::test_non_power_of_two_sizes shows it against a float64 DFT -- so parity for those
// sizes is stated against the DFT the reference's header specifies, and against SimdFFT for the sizes SimdFFT serves.)
#pragma once

namespace gr4b200 {
namespace {

__device__ __forceinline__ float2 cmulF(float2 a, float2 b) { return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }

// a[t][i] = x[t][i] * window[i] * conj(w[i]) for i < n, 0 for n <= i < m          (chirpConj holds conj(w))
template<bool RealInput>
__global__ void __launch_bounds__(256) bluesteinChirpIn(const float2* __restrict__ in, const float* __restrict__ inReal, const float* __restrict__ window, const float2* __restrict__ chirpConj, float2* __restrict__ a, int n, int m, long long batch) {
    const long long total = batch * m;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long t = idx / m;
        const int       i = static_cast<int>(idx - t * m);
        float2          v = make_float2(0.f, 0.f);
        if (i < n) {
            const float w = window != nullptr ? window[i] : 1.f;
            if constexpr (RealInput) {
                v = make_float2(__fmul_rn(inReal[t * n + i], w), 0.f);
            } else {
                const float2 x = in[t * n + i];
                v              = window != nullptr ? make_float2(__fmul_rn(x.x, w), __fmul_rn(x.y, w)) : x;
            }
            v = cmulF(v, chirpConj[i]);
        }
        a[idx] = v;
    }
}

// p[t][i] = conj(p[t][i] * chirpSpectrum[i])                                      (chirpSpectrum = FFT_m(b) / m)
__global__ void __launch_bounds__(256) bluesteinPointwise(float2* __restrict__ p, const float2* __restrict__ chirpSpectrum, int m, long long batch) {
    const long long total = batch * m;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float2 v = cmulF(p[idx], chirpSpectrum[idx % m]);
        p[idx]         = make_float2(v.x, -v.y);
    }
}

// X[t][k] = conj(c[t][k]) * conj(w[k]), k < n; realSpectrum: bins 0 and n/2 of a real signal's spectrum are real
__global__ void __launch_bounds__(256) bluesteinChirpOut(const float2* __restrict__ c, const float2* __restrict__ chirpConj, float2* __restrict__ out, int n, int m, long long batch, int realSpectrum) {
    const long long total = batch * n;
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long t = idx / n;
        const int       k = static_cast<int>(idx - t * n);
        const float2    v = c[t * m + k];
        float2          r = cmulF(make_float2(v.x, -v.y), chirpConj[k]);
        if (realSpectrum != 0 && (k == 0 || 2 * k == n)) {
            r.y = 0.f;
        }
        out[idx] = r;
    }
}

// the FFT block's planes from a spectrum in natural order (fft_common.hpp:22-56, 93-123; blocks/fourier fft.hpp:147-250).
// Complex input: four planes of n values {magnitude (fft-shifted), phase (fft-shifted), Re, Im}. Real input: four planes
// of n/2 values {magnitude and phase of bins [0, n/2), Re and Im of bins [n/2, n/2 + n/2)}. One thread per bin.
__global__ void __launch_bounds__(256) spectrumPlanesKernel(const float2* __restrict__ spectrum, float* __restrict__ signals, int n, long long batch, unsigned flags, int realInput) {
    const int       planeLen = realInput != 0 ? n / 2 : n;
    const long long total    = batch * planeLen;
    const bool      dB = (flags & GR4B200_FFT_OUTPUT_IN_DB) != 0, deg = (flags & GR4B200_FFT_OUTPUT_IN_DEG) != 0;
    const float     nf = static_cast<float>(n);
    for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total; idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long t   = idx / planeLen;
        const int       k   = static_cast<int>(idx - t * planeLen);
        const float2    x   = spectrum[t * n + k];
        const float     mag = __fdiv_rn(__fmul_rn(hypotf(x.x, x.y), 2.f), nf); // hypot * 2 / N in the reference's order
        const float     ph  = atan2f(x.y, x.x);
        float*          sig = signals + t * 4 * planeLen;
        if (realInput != 0) {
            const float2 upper   = spectrum[t * n + planeLen + k];
            sig[k]               = dB ? decibel(mag) : mag;
            sig[planeLen + k]    = deg ? toDegrees(ph) : ph;
            sig[2 * planeLen + k] = upper.x;
            sig[3 * planeLen + k] = upper.y;
        } else {
            int pos = k + (n - n / 2); // std::rotate(begin, begin + n/2, end): bin k lands at (k - n/2) mod n
            pos -= pos >= n ? n : 0;
            sig[pos]       = dB ? decibel(mag) : mag;
            sig[n + pos]   = deg ? toDegrees(ph) : ph;
            sig[2 * n + k] = x.x;
            sig[3 * n + k] = x.y;
        }
    }
}

inline int gridForElements(long long total) {
    const long long want = ceilDiv<long long>(total, 256);
    const long long cap  = static_cast<long long>(smCount()) * 16;
    return static_cast<int>(want < cap ? (want < 1 ? 1 : want) : cap);
}

// FFT of a length-m (power of two) sequence in double on the host: plan creation only
inline void hostFftDouble(std::vector<double>& re, std::vector<double>& im) {
    const size_t m = re.size();
    for (size_t i = 1, j = 0; i < m; ++i) {
        size_t bit = m >> 1;
        for (; j & bit; bit >>= 1) {
            j ^= bit;
        }
        j ^= bit;
        if (i < j) {
            std::swap(re[i], re[j]);
            std::swap(im[i], im[j]);
        }
    }
    for (size_t len = 2; len <= m; len <<= 1) {
        const double angle = -2.0 * 3.14159265358979323846 / static_cast<double>(len);
        for (size_t i = 0; i < m; i += len) {
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
}

inline bool bluesteinUpload(const void* host, size_t bytes, void** device) { return cudaMalloc(device, bytes) == cudaSuccess && cudaMemcpy(*device, host, bytes, cudaMemcpyHostToDevice) == cudaSuccess; }

} // namespace

constexpr size_t kBluesteinMaxN = 131072; // M = bit_ceil(2 N - 1) <= 262144, the largest power-of-two plan

// fills the Bluestein part of `plan` (plan->n set, not a size the radix kernels serve); false on allocation failure
inline bool bluesteinPlanCreate(gr4b200_fft_plan* plan, const float* window_host) {
    const size_t n = plan->n;
    size_t       m = 16;
    while (m < 2 * n - 1) {
        m <<= 1;
    }
    plan->bluesteinM = m;
    plan->inner      = gr4b200_fft_plan_create(m, nullptr);
    if (plan->inner == nullptr) {
        return false;
    }
    std::vector<float2> chirpConj(n);
    std::vector<double> bRe(m, 0.0), bIm(m, 0.0);
    for (size_t i = 0; i < n; ++i) {
        const unsigned long long sq    = static_cast<unsigned long long>(i) * i % (2 * n);
        const double             angle = 3.14159265358979323846 * static_cast<double>(sq) / static_cast<double>(n);
        const double             c = std::cos(angle), s = std::sin(angle);
        chirpConj[i]                   = make_float2(static_cast<float>(c), static_cast<float>(-s));
        bRe[i]                         = c;
        bIm[i]                         = s;
        if (i > 0) {
            bRe[m - i] = c;
            bIm[m - i] = s;
        }
    }
    hostFftDouble(bRe, bIm);
    std::vector<float2> spectrum(m);
    for (size_t i = 0; i < m; ++i) {
        spectrum[i] = make_float2(static_cast<float>(bRe[i] / static_cast<double>(m)), static_cast<float>(bIm[i] / static_cast<double>(m)));
    }
    bool ok = bluesteinUpload(chirpConj.data(), n * sizeof(float2), reinterpret_cast<void**>(&plan->chirpConj)) && bluesteinUpload(spectrum.data(), m * sizeof(float2), reinterpret_cast<void**>(&plan->chirpSpectrum));
    if (window_host != nullptr) {
        ok = ok && bluesteinUpload(window_host, n * sizeof(float), reinterpret_cast<void**>(&plan->windowN));
    }
    return ok;
}

// spectrum of `batch` transforms into `out` (n values each); in / inReal: complex or real input
inline int bluesteinSpectrum(gr4b200_fft_plan* plan, cudaStream_t stream, const float2* in, const float* inReal, float2* out, size_t batch) {
    const size_t n = plan->n, m = plan->bluesteinM;
    const size_t slice = std::max<size_t>(1, (size_t{1} << 25) / m); // transforms per pass: two work arrays of <= 256 MiB
    const size_t need  = 2 * std::min(batch, slice) * m;
    if (plan->workSize < need) {
        GR4B200_CUDA_TRY(cudaStreamSynchronize(stream)); // earlier launches may still use the old arrays
        cudaFree(plan->work);
        plan->work     = nullptr;
        plan->workSize = 0;
        GR4B200_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&plan->work), need * sizeof(float2)));
        plan->workSize = need;
    }
    for (size_t done = 0; done < batch; done += slice) {
        const size_t    count = std::min(slice, batch - done);
        const long long b     = static_cast<long long>(count);
        float2*         a     = plan->work;
        float2*         p     = plan->work + count * m;
        if (in != nullptr) {
            bluesteinChirpIn<false><<<gridForElements(b * static_cast<long long>(m)), 256, 0, stream>>>(in + done * n, nullptr, plan->windowN, plan->chirpConj, a, static_cast<int>(n), static_cast<int>(m), b);
        } else {
            bluesteinChirpIn<true><<<gridForElements(b * static_cast<long long>(m)), 256, 0, stream>>>(nullptr, inReal + done * n, plan->windowN, plan->chirpConj, a, static_cast<int>(n), static_cast<int>(m), b);
        }
        int status = checkLaunch("bluesteinChirpIn");
        status     = status == GR4B200_OK ? gr4b200_fft_c2c_cf32(plan->inner, stream, reinterpret_cast<const float*>(a), reinterpret_cast<float*>(p), count) : status;
        if (status != GR4B200_OK) {
            return status;
        }
        bluesteinPointwise<<<gridForElements(b * static_cast<long long>(m)), 256, 0, stream>>>(p, plan->chirpSpectrum, static_cast<int>(m), b);
        status = checkLaunch("bluesteinPointwise");
        status = status == GR4B200_OK ? gr4b200_fft_c2c_cf32(plan->inner, stream, reinterpret_cast<const float*>(p), reinterpret_cast<float*>(a), count) : status;
        if (status != GR4B200_OK) {
            return status;
        }
        bluesteinChirpOut<<<gridForElements(b * static_cast<long long>(n)), 256, 0, stream>>>(a, plan->chirpConj, out + done * n, static_cast<int>(n), static_cast<int>(m), b, in == nullptr ? 1 : 0);
        status = checkLaunch("bluesteinChirpOut");
        if (status != GR4B200_OK) {
            return status;
        }
    }
    return GR4B200_OK;
}

// the FFT block on a size served by Bluestein: spectrum into a plan-owned buffer, then the plane kernel and the
// optional post-passes (unwrapping, ranges) of the power-of-two path
inline int bluesteinBlock(gr4b200_fft_plan* plan, cudaStream_t stream, const float2* in, const float* inReal, size_t batch, unsigned flags, float* signals, float* ranges) {
    const size_t n = plan->n;
    if (inReal != nullptr && n % 2 != 0) {
        return fail("fft_block_f32: an odd fftSize has no half spectrum (N/2 bins); use an even size");
    }
    if (plan->spectrumSize < batch * n) {
        GR4B200_CUDA_TRY(cudaStreamSynchronize(stream));
        cudaFree(plan->spectrum);
        plan->spectrum     = nullptr;
        plan->spectrumSize = 0;
        GR4B200_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&plan->spectrum), batch * n * sizeof(float2)));
        plan->spectrumSize = batch * n;
    }
    int status = bluesteinSpectrum(plan, stream, in, inReal, plan->spectrum, batch);
    if (status != GR4B200_OK) {
        return status;
    }
    const bool      unwrap   = (flags & GR4B200_FFT_UNWRAP_PHASE) != 0;
    const int       realIn   = inReal != nullptr ? 1 : 0;
    const long long planeLen = static_cast<long long>(realIn ? n / 2 : n);
    spectrumPlanesKernel<<<gridForElements(static_cast<long long>(batch) * planeLen), 256, 0, stream>>>(plan->spectrum, signals, static_cast<int>(n), static_cast<long long>(batch), unwrap ? (flags & ~GR4B200_FFT_OUTPUT_IN_DEG) : flags, realIn);
    status = checkLaunch("spectrumPlanesKernel");
    if (status == GR4B200_OK && unwrap) {
        const int deg = (flags & GR4B200_FFT_OUTPUT_IN_DEG) != 0 ? 1 : 0;
        if (realIn) {
            unwrapHalfPlaneKernel<<<static_cast<int>(ceilDiv<size_t>(batch, 64)), 64, 0, stream>>>(signals, static_cast<long long>(batch), static_cast<int>(planeLen), deg);
        } else {
            unwrapPhaseKernel<<<static_cast<int>(ceilDiv<size_t>(batch, 64)), 64, 0, stream>>>(signals, static_cast<long long>(batch), static_cast<int>(n), deg);
        }
        status = checkLaunch("unwrapPhaseKernel");
    }
    if (status == GR4B200_OK && ranges != nullptr) {
        const long long rows = static_cast<long long>(batch) * 4;
        rangesKernel<<<static_cast<int>(ceilDiv<long long>(rows * 32, 256)), 256, 0, stream>>>(signals, ranges, rows, static_cast<int>(planeLen));
        status = checkLaunch("rangesKernel");
    }
    return status;
}

} // namespace gr4b200
