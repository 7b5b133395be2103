// Polyphase channelizer filter bank.
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace gr4b200 {
namespace {

// u[t][r] = sum_q h[r + q M] * x[(t - q) M + (M - 1 - r)], q ascending, products and sums rounded separately.
// xe = state ++ in with xe index = sample index + halo (halo = (P-1) M). One thread per output, consecutive threads take
// consecutive r => both the (reversed) sample reads and the stores are contiguous per warp.
__global__ void __launch_bounds__(256) pfbFilterKernel(const float2* __restrict__ in, const float2* __restrict__ state, const float* __restrict__ proto, float2* __restrict__ out, long long nFrames, int M, int P) {
    const long long total = nFrames * M;
    const long long halo  = static_cast<long long>(P - 1) * M;
    for (long long o = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long t = o / M;
        const int       r = static_cast<int>(o - t * M);
        float           accRe = 0.f, accIm = 0.f;
        for (int q = 0; q < P; ++q) {
            const long long idx = (t - q) * M + (M - 1 - r); // sample index, negative => history
            const float2    x   = idx >= 0 ? in[idx] : state[halo + idx];
            const float     h   = __ldg(proto + r + q * M);
            accRe               = __fadd_rn(accRe, __fmul_rn(h, x.x));
            accIm               = __fadd_rn(accIm, __fmul_rn(h, x.y));
        }
        out[o] = make_float2(accRe, accIm);
    }
}

__global__ void pfbUpdateState(const float2* __restrict__ oldState, const float2* __restrict__ in, float2* __restrict__ newState, long long halo, long long nIn) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < halo; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long q = nIn - halo + i;
        newState[i]       = q >= 0 ? in[q] : oldState[halo + q];
    }
}

} // namespace
} // namespace gr4b200

using namespace gr4b200;

struct gr4b200_pfb_plan {
    int     M        = 0;
    int     P        = 0;
    float*  proto    = nullptr;
    float2* state[2] = {nullptr, nullptr};
    int     current  = 0;
};

extern "C" {

gr4b200_pfb_plan* gr4b200_pfb_plan_create(const float* proto_host, size_t nChannels, size_t tapsPerBranch) {
    if (proto_host == nullptr || nChannels == 0 || tapsPerBranch == 0 || nChannels > (1u << 16) || tapsPerBranch > 4096) {
        fail("pfb_plan_create: bad arguments");
        return nullptr;
    }
    auto* plan = new gr4b200_pfb_plan;
    plan->M    = static_cast<int>(nChannels);
    plan->P    = static_cast<int>(tapsPerBranch);
    const size_t protoBytes = nChannels * tapsPerBranch * sizeof(float);
    const size_t haloBytes  = (tapsPerBranch - 1) * nChannels * sizeof(float2) + 16;
    bool         ok         = cudaMalloc(&plan->proto, protoBytes) == cudaSuccess && cudaMalloc(&plan->state[0], haloBytes) == cudaSuccess && cudaMalloc(&plan->state[1], haloBytes) == cudaSuccess;
    ok                      = ok && cudaMemcpy(plan->proto, proto_host, protoBytes, cudaMemcpyHostToDevice) == cudaSuccess;
    ok                      = ok && cudaMemset(plan->state[0], 0, haloBytes) == cudaSuccess && cudaMemset(plan->state[1], 0, haloBytes) == cudaSuccess;
    if (!ok) {
        checkCuda(cudaGetLastError(), "pfb_plan_create");
        gr4b200_pfb_plan_destroy(plan);
        return nullptr;
    }
    return plan;
}

int gr4b200_pfb_plan_destroy(gr4b200_pfb_plan* plan) {
    if (plan == nullptr) {
        return GR4B200_OK;
    }
    cudaFree(plan->proto);
    cudaFree(plan->state[0]);
    cudaFree(plan->state[1]);
    delete plan;
    return GR4B200_OK;
}

int gr4b200_pfb_plan_reset(gr4b200_pfb_plan* plan, void* stream) {
    if (plan == nullptr) {
        return fail("pfb_plan_reset: null plan");
    }
    const size_t haloBytes = static_cast<size_t>(plan->P - 1) * plan->M * sizeof(float2);
    return checkCuda(cudaMemsetAsync(plan->state[plan->current], 0, haloBytes, asStream(stream)), "pfb_plan_reset");
}

int gr4b200_pfb_filter_cf32(gr4b200_pfb_plan* plan, void* stream, const float* in, float* out, size_t nFrames) {
    if (plan == nullptr) {
        return fail("pfb_filter: null plan");
    }
    if (nFrames == 0) {
        return GR4B200_OK;
    }
    if (in == nullptr || out == nullptr) {
        return fail("pfb_filter: null buffer");
    }
    const auto      s     = asStream(stream);
    const long long total = static_cast<long long>(nFrames) * plan->M;
    const long long cap   = static_cast<long long>(smCount()) * 8;
    const long long want  = ceilDiv<long long>(total, 256);
    pfbFilterKernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, s>>>(reinterpret_cast<const float2*>(in), plan->state[plan->current], plan->proto, reinterpret_cast<float2*>(out), static_cast<long long>(nFrames), plan->M, plan->P);
    const long long halo = static_cast<long long>(plan->P - 1) * plan->M;
    if (halo > 0) {
        pfbUpdateState<<<static_cast<int>(std::min<long long>(ceilDiv<long long>(halo, 256), cap)), 256, 0, s>>>(plan->state[plan->current], reinterpret_cast<const float2*>(in), plan->state[plan->current ^ 1], halo, total);
        plan->current ^= 1;
    }
    return checkLaunch("pfbFilterKernel");
}

} // extern "C"
