// Polyphase rational resampler (interpolate by L, FIR, decimate by M) on complex<float> streams (sm_100a).
// There is no such block in the reference: own definition, stated in DESIGN.md (3.5):
//   y[m] = sum_{k<P} h[p_m + k L] * x[q_m - k],  p_m = (m M) mod L,  q_m = floor(m M / L),  P = ceil(K / L),
//   k ascending, acc = fma(h, x, acc); the zero-stuffed samples of the textbook form are never touched.
// Kernel: persistent CTAs over tiles of consecutive outputs. A tile needs one contiguous stretch of inputs
// (tileOut * M / L + P samples): it is staged in shared memory with coalesced 8-byte loads, the phase-major tap table
// hT[p][k] = h[p + k L] sits in shared memory too (when it fits), every thread forms outputs m0 + tid, m0 + tid +
// Threads, ...: stores are coalesced, (p, q) advance incrementally (no 64-bit division in the inner loop).
// HBM traffic: 8 B per input sample + 8 B per output sample.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "fir_core.cuh"

namespace gr4b200 {
namespace {

constexpr int kResamplerThreads = 256;
constexpr int kResamplerTileIn  = 4096; // staged input samples per tile (32 KB)

struct ResamplerArgs {
    const float2* in;
    const float2* state; // P-1 samples in front of in[0]
    const float*  tapsT; // [L][rowPitch], zero padded
    float2*       out;
    long long     nIn, nOut, nTiles;
    int           L, M, P;
    int           rowPitch;     // floats per tap row: P rounded up to a multiple of 4 (16-byte row loads)
    int           tileOut;      // outputs per tile
    int           tapsInShared; // 1: the tap table fits next to the sample tile
};

__global__ void __launch_bounds__(kResamplerThreads) resamplerKernel(ResamplerArgs a) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float2* sX    = reinterpret_cast<float2*>(smemRaw);
    float*  sTaps = reinterpret_cast<float*>(sX + kResamplerTileIn);
    const int tid = threadIdx.x;
    const int L = a.L, M = a.M, P = a.P;
    const float* taps = a.tapsT;
    if (a.tapsInShared != 0) {
        for (int i = tid; i < L * a.rowPitch; i += kResamplerThreads) {
            sTaps[i] = __ldg(a.tapsT + i);
        }
        taps = sTaps;
    }
    const long long halo  = P - 1;
    const int       stepP = (kResamplerThreads * M) % L; // advancing m by Threads moves (p, q) by (stepP, stepQ) (+ carry)
    const int       stepQ = (kResamplerThreads * M) / L;
    for (long long tile = blockIdx.x; tile < a.nTiles; tile += gridDim.x) {
        const long long m0 = tile * a.tileOut;
        const long long m1 = m0 + a.tileOut < a.nOut ? m0 + a.tileOut : a.nOut;
        const long long qLo = (m0 * M) / L - halo;          // first input sample the tile touches (may be negative: history)
        const long long qHi = ((m1 - 1) * M) / L;           // last one
        const int       count = static_cast<int>(qHi - qLo + 1);
        __syncthreads(); // the previous tile's readers are done (and the tap table is complete on the first pass)
        for (int i = tid; i < count; i += kResamplerThreads) {
            const long long q = qLo + i;
            sX[i]             = q >= 0 ? ldStream2(a.in + q) : __ldg(a.state + halo + q);
        }
        __syncthreads();
        long long m = m0 + tid;
        if (m < m1) {
            const unsigned long long mm = static_cast<unsigned long long>(m) * static_cast<unsigned long long>(M);
            int                      p  = static_cast<int>(mm % static_cast<unsigned long long>(L));
            int                      q  = static_cast<int>(static_cast<long long>(mm / static_cast<unsigned long long>(L)) - qLo); // index into sX
            for (; m < m1; m += kResamplerThreads) {
                const float4* h   = reinterpret_cast<const float4*>(taps + p * a.rowPitch);
                const float2* x   = sX + q;
                Packed        acc = packPair(0.f, 0.f);
                for (int k4 = 0; k4 < P / 4; ++k4) { // four taps per 16-byte load
                    const float4 t  = h[k4];
                    const float2 v0 = x[-4 * k4], v1 = x[-4 * k4 - 1], v2 = x[-4 * k4 - 2], v3 = x[-4 * k4 - 3];
                    acc             = fmaV(t.x, packPair(v0.x, v0.y), acc);
                    acc             = fmaV(t.y, packPair(v1.x, v1.y), acc);
                    acc             = fmaV(t.z, packPair(v2.x, v2.y), acc);
                    acc             = fmaV(t.w, packPair(v3.x, v3.y), acc);
                }
                for (int k = P / 4 * 4; k < P; ++k) { // P % 4 remaining taps (never the zero padding: 0 * inf would differ)
                    const float2 v = x[-k];
                    acc            = fmaV(taps[p * a.rowPitch + k], packPair(v.x, v.y), acc);
                }
                stStream2(a.out + m, make_float2(packedLo(acc), packedHi(acc)));
                p += stepP;
                q += stepQ;
                if (p >= L) {
                    p -= L;
                    q += 1;
                }
            }
        }
    }
}

__global__ void resamplerUpdateState(const float2* __restrict__ oldState, const float2* __restrict__ in, float2* __restrict__ newState, long long halo, long long nIn) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < halo; i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long q = nIn - halo + i;
        newState[i]       = q >= 0 ? in[q] : oldState[halo + q];
    }
}

} // namespace
} // namespace gr4b200

using namespace gr4b200;

struct gr4b200_resampler_plan {
    int     device = 0; // the device the plan's memory lives on
    int     L = 1, M = 1, P = 1;
    int     rowPitch = 4;       // P rounded up to a multiple of 4
    float*  tapsT    = nullptr; // device, [L][rowPitch]
    float2* state[2] = {nullptr, nullptr};
    int     current  = 0;
};

extern "C" {

gr4b200_resampler_plan* gr4b200_resampler_plan_create(const float* taps_host, size_t nTaps, size_t interpolation, size_t decimation) {
    if (taps_host == nullptr || nTaps == 0 || interpolation == 0 || decimation == 0 || interpolation > (1u << 16) || decimation > (1u << 16) || nTaps > (1u << 22)) {
        fail("resampler_plan_create: need taps and 1 <= interpolation, decimation <= 65536");
        return nullptr;
    }
    auto* plan   = new gr4b200_resampler_plan;
    plan->device = currentDevice();
    plan->L      = static_cast<int>(interpolation);
    plan->M    = static_cast<int>(decimation);
    plan->P    = static_cast<int>((nTaps + interpolation - 1) / interpolation);
    if (plan->P + 2 * plan->M / plan->L + 2 > kResamplerTileIn / 2) {
        fail("resampler_plan_create: too many taps per phase for the shared-memory tile");
        delete plan;
        return nullptr;
    }
    plan->rowPitch = (plan->P + 3) / 4 * 4;
    std::vector<float> table(static_cast<size_t>(plan->L) * plan->rowPitch, 0.f);
    for (int p = 0; p < plan->L; ++p) {
        for (int k = 0; k < plan->P; ++k) {
            const size_t index = static_cast<size_t>(p) + static_cast<size_t>(k) * plan->L;
            table[static_cast<size_t>(p) * plan->rowPitch + k] = index < nTaps ? taps_host[index] : 0.f;
        }
    }
    const size_t haloBytes = static_cast<size_t>(plan->P) * sizeof(float2);
    bool         ok        = cudaMalloc(&plan->tapsT, table.size() * sizeof(float)) == cudaSuccess && cudaMemcpy(plan->tapsT, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
    ok                     = ok && cudaMalloc(&plan->state[0], haloBytes) == cudaSuccess && cudaMalloc(&plan->state[1], haloBytes) == cudaSuccess;
    ok                     = ok && cudaMemset(plan->state[0], 0, haloBytes) == cudaSuccess && cudaMemset(plan->state[1], 0, haloBytes) == cudaSuccess;
    if (!ok) {
        checkCuda(cudaGetLastError(), "resampler_plan_create");
        gr4b200_resampler_plan_destroy(plan);
        return nullptr;
    }
    return plan;
}

int gr4b200_resampler_plan_destroy(gr4b200_resampler_plan* plan) {
    if (plan == nullptr) {
        return GR4B200_OK;
    }
    cudaFree(plan->tapsT);
    cudaFree(plan->state[0]);
    cudaFree(plan->state[1]);
    delete plan;
    return GR4B200_OK;
}

int gr4b200_resampler_plan_reset(gr4b200_resampler_plan* plan, void* stream) {
    if (plan == nullptr) {
        return fail("resampler_plan_reset: null plan");
    }
    return checkCuda(cudaMemsetAsync(plan->state[plan->current], 0, static_cast<size_t>(plan->P) * sizeof(float2), asStream(stream)), "resampler_plan_reset");
}

int gr4b200_resampler_cf32(gr4b200_resampler_plan* plan, void* stream, const float* in, float* out, size_t nIn) {
    if (plan == nullptr) {
        return fail("resampler: null plan");
    }
    if (const int status = checkPlanDevice(plan->device, "resampler"); status != GR4B200_OK) {
        return status;
    }
    if (nIn % static_cast<size_t>(plan->M) != 0) {
        return fail("resampler: nIn must be a multiple of the decimation factor", GR4B200_INSUFFICIENT_INPUT_ITEMS);
    }
    if (nIn == 0) {
        return GR4B200_OK;
    }
    if (in == nullptr || out == nullptr || reinterpret_cast<uintptr_t>(in) % 8 != 0 || reinterpret_cast<uintptr_t>(out) % 8 != 0) {
        return fail("resampler: null or misaligned buffer");
    }
    ResamplerArgs a{};
    a.in    = reinterpret_cast<const float2*>(in);
    a.state = plan->state[plan->current];
    a.tapsT = plan->tapsT;
    a.out   = reinterpret_cast<float2*>(out);
    a.nIn   = static_cast<long long>(nIn);
    a.nOut  = a.nIn / plan->M * plan->L;
    a.L = plan->L, a.M = plan->M, a.P = plan->P;
    a.rowPitch = plan->rowPitch;
    // outputs per tile such that the inputs it touches fit the staged tile: tileOut * M / L + P + 1 <= kResamplerTileIn
    long long tileOut = (static_cast<long long>(kResamplerTileIn - plan->P - 2) * plan->L) / plan->M;
    tileOut           = std::max<long long>(1, std::min<long long>(tileOut, 16384));
    a.tileOut         = static_cast<int>(tileOut);
    a.nTiles          = ceilDiv<long long>(a.nOut, tileOut);
    const size_t tapBytes = static_cast<size_t>(plan->L) * plan->rowPitch * sizeof(float);
    a.tapsInShared        = tapBytes <= 64 * 1024 ? 1 : 0;
    const size_t smem     = kResamplerTileIn * sizeof(float2) + (a.tapsInShared != 0 ? tapBytes : 0);
    GR4B200_CUDA_TRY(cudaFuncSetAttribute(resamplerKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int ctasPerSm = 0;
    GR4B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, resamplerKernel, kResamplerThreads, smem));
    static const int gridMult = [] { const char* e = std::getenv("GR4B200_RESAMPLER_GRID_MULT"); return e != nullptr ? std::atoi(e) : 0; }(); // 0: one CTA per tile (+10-15 % over a resident grid)
    const long long  cap      = gridMult > 0 ? static_cast<long long>(smCount()) * (ctasPerSm < 1 ? 1 : ctasPerSm) * gridMult : a.nTiles;
    const auto      s    = asStream(stream);
    resamplerKernel<<<static_cast<int>(a.nTiles < cap ? a.nTiles : cap), kResamplerThreads, smem, s>>>(a);
    const long long halo = plan->P - 1;
    if (halo > 0) {
        resamplerUpdateState<<<static_cast<int>(ceilDiv<long long>(halo, 256)), 256, 0, s>>>(plan->state[plan->current], a.in, plan->state[plan->current ^ 1], halo, a.nIn);
        plan->current ^= 1;
    }
    return checkLaunch("resamplerKernel", halo > 0 ? 2u : 1u);
}

} // extern "C"
