// Sample-format converters on either side of the path: gr::blocks::type::converter::InterleavedToComplex<R, complex<float>>
// and ComplexToInterleaved<complex<float>, R> for R in {float, int16, int8}
// (blocks/basic/include/gnuradio-4.0/basic/ConverterBlocks.hpp:233-277) -- what an SDR source or a file of interleaved
// I/Q integers needs in front of the FIR, and a fixed-point sink behind it. Copying int16 I/Q over PCIe and widening it
// on the device halves the host->device bytes per sample (int8: a quarter).
//
// Arithmetic: R -> float is the exact static_cast; float -> integer is static_cast, i.e. truncation toward zero. For
// values outside the integer type the C++ cast is undefined; this follows what the compiled reference does on x86-64
// (cvttss2si to a 32-bit integer -- 0x80000000 for NaN and |x| >= 2^31 -- then the low 8 / 16 bits).
// HBM-bound: one thread per PAIR of samples (16 bytes of floats, 8 / 4 bytes of integers), both sides coalesced,
// four pairs in flight per thread, one CTA per block of work.
#include <climits>
#include <cstdint>

#include "common.cuh"

namespace gr4b200 {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll  = 4;

__device__ __forceinline__ int truncateLikeX86(float x) { return fabsf(x) < 2147483648.f ? __float2int_rz(x) : INT_MIN; }

template<typename R>
struct Packed; // two complex samples = four items of R
template<>
struct Packed<int16_t> {
    using type = short4;
    __device__ static float4 widen(short4 v) { return make_float4(static_cast<float>(v.x), static_cast<float>(v.y), static_cast<float>(v.z), static_cast<float>(v.w)); }
    __device__ static short4 narrow(float4 v) { return make_short4(static_cast<short>(truncateLikeX86(v.x)), static_cast<short>(truncateLikeX86(v.y)), static_cast<short>(truncateLikeX86(v.z)), static_cast<short>(truncateLikeX86(v.w))); }
};
template<>
struct Packed<int8_t> {
    using type = char4;
    __device__ static float4 widen(char4 v) { return make_float4(static_cast<float>(v.x), static_cast<float>(v.y), static_cast<float>(v.z), static_cast<float>(v.w)); }
    __device__ static char4 narrow(float4 v) { return make_char4(static_cast<signed char>(truncateLikeX86(v.x)), static_cast<signed char>(truncateLikeX86(v.y)), static_cast<signed char>(truncateLikeX86(v.z)), static_cast<signed char>(truncateLikeX86(v.w))); }
};

template<typename R>
__global__ void __launch_bounds__(kThreads) widenPairsKernel(const typename Packed<R>::type* __restrict__ in, float4* __restrict__ out, size_t pairs) {
    const size_t base = static_cast<size_t>(blockIdx.x) * (kThreads * kUnroll) + threadIdx.x;
    typename Packed<R>::type v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const size_t i = base + static_cast<size_t>(u) * kThreads;
        if (i < pairs) {
            v[u] = __ldg(in + i);
        }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const size_t i = base + static_cast<size_t>(u) * kThreads;
        if (i < pairs) {
            stStream4(out + i, Packed<R>::widen(v[u]));
        }
    }
}

template<typename R>
__global__ void __launch_bounds__(kThreads) narrowPairsKernel(const float4* __restrict__ in, typename Packed<R>::type* __restrict__ out, size_t pairs) {
    const size_t base = static_cast<size_t>(blockIdx.x) * (kThreads * kUnroll) + threadIdx.x;
    float4       v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const size_t i = base + static_cast<size_t>(u) * kThreads;
        if (i < pairs) {
            v[u] = ldStream4(in + i);
        }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
        const size_t i = base + static_cast<size_t>(u) * kThreads;
        if (i < pairs) {
            out[i] = Packed<R>::narrow(v[u]);
        }
    }
}

// item by item: the odd last sample, and buffers that are not aligned for the packed accesses
template<typename R>
__global__ void __launch_bounds__(kThreads) widenItemsKernel(const R* __restrict__ in, float* __restrict__ out, size_t items) {
    const size_t stride = static_cast<size_t>(gridDim.x) * kThreads;
    for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < items; i += stride) {
        out[i] = static_cast<float>(in[i]);
    }
}
template<typename R>
__global__ void __launch_bounds__(kThreads) narrowItemsKernel(const float* __restrict__ in, R* __restrict__ out, size_t items) {
    const size_t stride = static_cast<size_t>(gridDim.x) * kThreads;
    for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < items; i += stride) {
        out[i] = static_cast<R>(truncateLikeX86(in[i]));
    }
}

int itemGrid(size_t items) {
    const size_t blocks = ceilDiv<size_t>(items, kThreads);
    const size_t cap    = static_cast<size_t>(smCount()) * 16;
    return static_cast<int>(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

template<typename R>
int widen(cudaStream_t stream, const void* in, float* out, size_t nComplex) {
    using P              = typename Packed<R>::type;
    const bool   aligned = reinterpret_cast<uintptr_t>(in) % sizeof(P) == 0 && reinterpret_cast<uintptr_t>(out) % 16 == 0;
    const size_t pairs   = aligned ? nComplex / 2 : 0;
    if (pairs > 0) {
        widenPairsKernel<R><<<static_cast<unsigned>(ceilDiv<size_t>(pairs, kThreads * kUnroll)), kThreads, 0, stream>>>(static_cast<const P*>(in), reinterpret_cast<float4*>(out), pairs);
    }
    const size_t rest = 2 * nComplex - 4 * pairs; // items
    if (rest > 0) {
        widenItemsKernel<R><<<itemGrid(rest), kThreads, 0, stream>>>(static_cast<const R*>(in) + 4 * pairs, out + 4 * pairs, rest);
    }
    return checkLaunch("widenPairsKernel", (pairs > 0 ? 1u : 0u) + (rest > 0 ? 1u : 0u));
}

template<typename R>
int narrow(cudaStream_t stream, const float* in, void* out, size_t nComplex) {
    using P              = typename Packed<R>::type;
    const bool   aligned = reinterpret_cast<uintptr_t>(out) % sizeof(P) == 0 && reinterpret_cast<uintptr_t>(in) % 16 == 0;
    const size_t pairs   = aligned ? nComplex / 2 : 0;
    if (pairs > 0) {
        narrowPairsKernel<R><<<static_cast<unsigned>(ceilDiv<size_t>(pairs, kThreads * kUnroll)), kThreads, 0, stream>>>(reinterpret_cast<const float4*>(in), static_cast<P*>(out), pairs);
    }
    const size_t rest = 2 * nComplex - 4 * pairs;
    if (rest > 0) {
        narrowItemsKernel<R><<<itemGrid(rest), kThreads, 0, stream>>>(in + 4 * pairs, static_cast<R*>(out) + 4 * pairs, rest);
    }
    return checkLaunch("narrowPairsKernel", (pairs > 0 ? 1u : 0u) + (rest > 0 ? 1u : 0u));
}

} // namespace
} // namespace gr4b200

using namespace gr4b200;

extern "C" {

int gr4b200_interleaved_to_complex_cf32(void* stream, int item_type, const void* interleaved, float* out, size_t n_complex) {
    if (n_complex == 0) {
        return GR4B200_OK;
    }
    if (interleaved == nullptr || out == nullptr || reinterpret_cast<uintptr_t>(out) % 4 != 0) {
        return fail("interleaved_to_complex: null or misaligned buffer");
    }
    if (n_complex > (size_t{1} << 40)) {
        return fail("interleaved_to_complex: too many samples for one call");
    }
    switch (item_type) {
    case GR4B200_ITEM_F32: // (re, im) pairs of floats ARE complex<float>: a copy
        if (reinterpret_cast<uintptr_t>(interleaved) % 4 != 0) {
            return fail("interleaved_to_complex: misaligned float input");
        }
        return checkCuda(cudaMemcpyAsync(out, interleaved, n_complex * 2 * sizeof(float), cudaMemcpyDeviceToDevice, asStream(stream)), "interleaved_to_complex copy");
    case GR4B200_ITEM_I16:
        if (reinterpret_cast<uintptr_t>(interleaved) % 2 != 0) {
            return fail("interleaved_to_complex: misaligned int16 input");
        }
        return widen<int16_t>(asStream(stream), interleaved, out, n_complex);
    case GR4B200_ITEM_I8: return widen<int8_t>(asStream(stream), interleaved, out, n_complex);
    default: return fail("interleaved_to_complex: unknown item type");
    }
}

int gr4b200_complex_to_interleaved_cf32(void* stream, int item_type, const float* in, void* interleaved, size_t n_complex) {
    if (n_complex == 0) {
        return GR4B200_OK;
    }
    if (interleaved == nullptr || in == nullptr || reinterpret_cast<uintptr_t>(in) % 4 != 0) {
        return fail("complex_to_interleaved: null or misaligned buffer");
    }
    if (n_complex > (size_t{1} << 40)) {
        return fail("complex_to_interleaved: too many samples for one call");
    }
    switch (item_type) {
    case GR4B200_ITEM_F32:
        if (reinterpret_cast<uintptr_t>(interleaved) % 4 != 0) {
            return fail("complex_to_interleaved: misaligned float output");
        }
        return checkCuda(cudaMemcpyAsync(interleaved, in, n_complex * 2 * sizeof(float), cudaMemcpyDeviceToDevice, asStream(stream)), "complex_to_interleaved copy");
    case GR4B200_ITEM_I16:
        if (reinterpret_cast<uintptr_t>(interleaved) % 2 != 0) {
            return fail("complex_to_interleaved: misaligned int16 output");
        }
        return narrow<int16_t>(asStream(stream), in, interleaved, n_complex);
    case GR4B200_ITEM_I8: return narrow<int8_t>(asStream(stream), in, interleaved, n_complex);
    default: return fail("complex_to_interleaved: unknown item type");
    }
}

} // extern "C"
