// Elementwise complex<float> math: MathOpImpl / MathOpMultiPortImpl / Decimator device bodies.
// HBM-bound streaming kernels: 16-byte vector accesses, 4 independent loads in flight per thread, one CTA per 16 KB block.
#include "common.cuh"

namespace gr4b200 {
namespace {

constexpr int kThreads   = 256;
// Grid: ONE CTA per block of work, no persistent grid-stride loop. For pure streaming kernels the hardware CTA scheduler
// beats a resident grid that strides through memory in lockstep: 424-429 GS/s against 367-376 GS/s at 2^28 samples
// (profiles/r01y_time_mathop_variants.jsonl; 8, 4 and 16 resident CTAs per SM: 90 %, 90 %, 94 % of the copy peak).
#ifndef GR4B200_MATHOP_CTAS
#define GR4B200_MATHOP_CTAS 0
#endif
constexpr int kCtasPerSm = GR4B200_MATHOP_CTAS; // 0: no cap, one CTA per kThreads * kUnroll vectors
constexpr int kUnroll    = 4;
#ifdef GR4B200_MATHOP_PLAIN
__device__ __forceinline__ float4 ldVec(const float4* p) { return *p; }
__device__ __forceinline__ void   stVec(float4* p, float4 v) { *p = v; }
#else
__device__ __forceinline__ float4 ldVec(const float4* p) { return ldStream4(p); }
__device__ __forceinline__ void   stVec(float4* p, float4 v) { stStream4(p, v); }
#endif

template<int Op>
__device__ __forceinline__ float2 applyOp(float2 a, float2 v) {
    if constexpr (Op == GR4B200_OP_ADD) {
        return make_float2(__fadd_rn(a.x, v.x), __fadd_rn(a.y, v.y));
    } else if constexpr (Op == GR4B200_OP_SUBTRACT) {
        return make_float2(__fsub_rn(a.x, v.x), __fsub_rn(a.y, v.y));
    } else if constexpr (Op == GR4B200_OP_MULTIPLY) {
        return complexMulAnnexG(a.x, a.y, v.x, v.y);
    } else {
        return complexDivAnnexG(a.x, a.y, v.x, v.y);
    }
}

// out[i] = in[i] op value; n2 = number of float4 (sample pairs)
template<int Op>
__global__ void __launch_bounds__(kThreads) mathopConstVec4(const float4* __restrict__ in, float4* __restrict__ out, size_t n2, float2 value) {
    const size_t stride = static_cast<size_t>(gridDim.x) * kThreads;
    size_t       i      = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x;
    for (; i + (kUnroll - 1) * stride < n2; i += kUnroll * stride) {
        float4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            v[u] = ldVec(in + i + u * stride);
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const float2 lo = applyOp<Op>(make_float2(v[u].x, v[u].y), value);
            const float2 hi = applyOp<Op>(make_float2(v[u].z, v[u].w), value);
            stVec(out + i + u * stride, make_float4(lo.x, lo.y, hi.x, hi.y));
        }
    }
    for (; i < n2; i += stride) {
        const float4 v  = ldVec(in + i);
        const float2 lo = applyOp<Op>(make_float2(v.x, v.y), value);
        const float2 hi = applyOp<Op>(make_float2(v.z, v.w), value);
        stVec(out + i, make_float4(lo.x, lo.y, hi.x, hi.y));
    }
}

template<int Op>
__global__ void __launch_bounds__(kThreads) mathopConstVec2(const float2* __restrict__ in, float2* __restrict__ out, size_t n, float2 value) {
    const size_t stride = static_cast<size_t>(gridDim.x) * kThreads;
    for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < n; i += stride) {
        stStream2(out + i, applyOp<Op>(ldStream2(in + i), value));
    }
}

struct MultiInputs {
    const float2* in[32];
};

// left fold over the inputs, one pass: out = ((in0 op in1) op in2) ...  (the reference makes nInputs passes)
template<int Op>
__global__ void __launch_bounds__(kThreads) mathopMulti(MultiInputs inputs, int nInputs, float2* __restrict__ out, size_t n) {
    const size_t stride = static_cast<size_t>(gridDim.x) * kThreads;
    for (size_t i = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; i < n; i += stride) {
        float2 acc = ldStream2(inputs.in[0] + i);
        for (int k = 1; k < nInputs; ++k) {
            acc = applyOp<Op>(acc, ldStream2(inputs.in[k] + i));
        }
        stStream2(out + i, acc);
    }
}

__global__ void __launch_bounds__(kThreads) decimateKernel(const float2* __restrict__ in, float2* __restrict__ out, size_t nOut, size_t decim) {
    const size_t stride = static_cast<size_t>(gridDim.x) * kThreads;
    for (size_t j = static_cast<size_t>(blockIdx.x) * kThreads + threadIdx.x; j < nOut; j += stride) {
        out[j] = in[j * decim];
    }
}

int gridFor(size_t items) {
    const size_t wanted = ceilDiv<size_t>(items, kThreads);
    const size_t cap    = kCtasPerSm > 0 ? static_cast<size_t>(smCount()) * kCtasPerSm : wanted;
    return static_cast<int>(wanted < cap ? (wanted == 0 ? 1 : wanted) : cap);
}

template<int Op>
int launchConst(cudaStream_t stream, const float* in, float* out, size_t n, float2 value) {
    const bool aligned16 = (reinterpret_cast<uintptr_t>(in) % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
    if (aligned16) {
        const size_t n2 = n / 2;
        if (n2 > 0) {
            mathopConstVec4<Op><<<gridFor(ceilDiv<size_t>(n2, kUnroll)), kThreads, 0, stream>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), n2, value);
        }
        if (n % 2 != 0) {
            mathopConstVec2<Op><<<1, 32, 0, stream>>>(reinterpret_cast<const float2*>(in) + (n - 1), reinterpret_cast<float2*>(out) + (n - 1), 1, value);
        }
    } else {
        mathopConstVec2<Op><<<gridFor(n), kThreads, 0, stream>>>(reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out), n, value);
    }
    return checkLaunch("mathopConst", aligned16 ? (n / 2 > 0 ? 1u : 0u) + (n % 2 != 0 ? 1u : 0u) : 1u);
}

} // namespace
} // namespace gr4b200

using namespace gr4b200;

extern "C" {

int gr4b200_mathop_const_cf32(void* stream, int op, const float* in, float* out, size_t n, float valueRe, float valueIm) {
    if (n == 0) {
        return GR4B200_OK;
    }
    if (in == nullptr || out == nullptr) {
        return fail("mathop_const: null buffer");
    }
    if (reinterpret_cast<uintptr_t>(in) % 8 != 0 || reinterpret_cast<uintptr_t>(out) % 8 != 0) {
        return fail("mathop_const: complex<float> buffers must be 8-byte aligned");
    }
    const float2 value = make_float2(valueRe, valueIm);
    switch (op) {
    case GR4B200_OP_ADD: return launchConst<GR4B200_OP_ADD>(asStream(stream), in, out, n, value);
    case GR4B200_OP_SUBTRACT: return launchConst<GR4B200_OP_SUBTRACT>(asStream(stream), in, out, n, value);
    case GR4B200_OP_MULTIPLY: return launchConst<GR4B200_OP_MULTIPLY>(asStream(stream), in, out, n, value);
    case GR4B200_OP_DIVIDE: return launchConst<GR4B200_OP_DIVIDE>(asStream(stream), in, out, n, value);
    default: return fail("mathop_const: unknown op");
    }
}

int gr4b200_mathop_multi_cf32(void* stream, int op, const float* const* ins_host, size_t nInputs, float* out, size_t n) {
    if (nInputs < 1 || nInputs > 32) { // Math.hpp:93 Limits<1U, 32U>
        return fail("mathop_multi: n_inputs must be in [1, 32]");
    }
    if (n == 0) {
        return GR4B200_OK;
    }
    MultiInputs inputs{};
    for (size_t k = 0; k < nInputs; ++k) {
        if (ins_host[k] == nullptr || reinterpret_cast<uintptr_t>(ins_host[k]) % 8 != 0) {
            return fail("mathop_multi: null or misaligned input");
        }
        inputs.in[k] = reinterpret_cast<const float2*>(ins_host[k]);
    }
    if (out == nullptr || reinterpret_cast<uintptr_t>(out) % 8 != 0) {
        return fail("mathop_multi: null or misaligned output");
    }
    float2*    o    = reinterpret_cast<float2*>(out);
    const int  grid = gridFor(n);
    const auto s    = asStream(stream);
    switch (op) {
    case GR4B200_OP_ADD: mathopMulti<GR4B200_OP_ADD><<<grid, kThreads, 0, s>>>(inputs, static_cast<int>(nInputs), o, n); break;
    case GR4B200_OP_SUBTRACT: mathopMulti<GR4B200_OP_SUBTRACT><<<grid, kThreads, 0, s>>>(inputs, static_cast<int>(nInputs), o, n); break;
    case GR4B200_OP_MULTIPLY: mathopMulti<GR4B200_OP_MULTIPLY><<<grid, kThreads, 0, s>>>(inputs, static_cast<int>(nInputs), o, n); break;
    case GR4B200_OP_DIVIDE: mathopMulti<GR4B200_OP_DIVIDE><<<grid, kThreads, 0, s>>>(inputs, static_cast<int>(nInputs), o, n); break;
    default: return fail("mathop_multi: unknown op");
    }
    return checkLaunch("mathopMulti");
}

int gr4b200_decimate_cf32(void* stream, const float* in, float* out, size_t nIn, size_t decim) {
    if (decim == 0) {
        return fail("decimate: decim must be >= 1");
    }
    const size_t nOut = ceilDiv<size_t>(nIn, decim);
    if (nOut == 0) {
        return GR4B200_OK;
    }
    decimateKernel<<<gridFor(nOut), kThreads, 0, asStream(stream)>>>(reinterpret_cast<const float2*>(in), reinterpret_cast<float2*>(out), nOut, decim);
    return checkLaunch("decimate");
}

} // extern "C"
