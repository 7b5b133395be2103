// Polyphase channelizer filter bank.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "fft_radix.cuh"
#include "fir_core.cuh"

namespace gr4b200 {
namespace {

// u[t][r] = sum_q h[r + q M] * x[(t - q) M + (M - 1 - r)], q ascending, acc = fma(h, x, acc) (our own definition: there is
// no reference implementation; one rounding per tap halves the fp32-pipe time of this memory-bound kernel).
// Branch r sees every M-th sample: with sample(t) = x[t M + M - 1 - r] the output is a P-tap FIR over sample(t), so a
// thread that owns one branch and walks along t needs ONE new 8-byte load per output. The last samples live in a
// register ring of P + 4 slots addressed at compile time (the frame loop is unrolled ring-size-fold): four outputs are
// formed together (four independent accumulation chains), and the four slots they refill are exactly the ones none of
// them reads. New samples are loaded up to one full ring ahead into a second register set, so every thread keeps that
// many loads in flight while it computes. The P taps of the branch sit in registers.
// Consecutive threads own consecutive branches: the (reversed) loads and the stores of a warp are contiguous.
// grid.x = stretches of frames, grid.y * blockDim.x covers the branches; a stretch re-reads P-1 frames of history.
// Arithmetic: packed f32x2 fused multiply-add (fir_core.cuh fmaV), one FFMA2 per tap and complex sample.
template<int P>
__global__ void __launch_bounds__(256, 2) pfbStreamKernel(const float2* __restrict__ in, const float2* __restrict__ state, const float* __restrict__ proto, float2* __restrict__ out, long long nFrames, int M, long long framesPerStretch) {
    constexpr int Ring  = P + 4;
    constexpr int Ahead = P >= 16 ? Ring / 2 : Ring; // look-ahead depth in frames (register budget: 128 per thread)
    static_assert(Ring % Ahead == 0 && Ring % 4 == 0, "slots must be compile-time constants across ring turns");
    const int     r    = blockIdx.y * blockDim.x + threadIdx.x;
    if (r >= M) {
        return;
    }
    const long long      halo = static_cast<long long>(P - 1) * M;
    const long long      t0   = static_cast<long long>(blockIdx.x) * framesPerStretch;
    const long long      t1   = t0 + framesPerStretch < nFrames ? t0 + framesPerStretch : nFrames;
    const int            col  = M - 1 - r;
    float                h[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        h[q] = __ldg(proto + r + static_cast<long long>(q) * M);
    }
    auto sample = [&](long long t) -> Packed { // frame t of this branch; negative frames come from the carried history
        const long long idx = t * M + col;
        const float2    v   = idx >= 0 ? ldStream2(in + idx) : __ldg(state + halo + idx);
        return packPair(v.x, v.y);
    };
    const Packed zero = packPair(0.f, 0.f);
    Packed       ring[Ring]; // frame tb + j lives in slot j: sample(tb + j - q) = ring[(j - q) mod Ring]
    Packed       ahead[Ahead]; // frame f waits in slot f mod Ahead, loaded Ahead frames before it is needed
#pragma unroll
    for (int j = 0; j < Ring; ++j) {
        ring[j] = zero;
    }
#pragma unroll
    for (int i = 0; i + 1 < P; ++i) {
        ring[Ring - 1 - i] = sample(t0 - 1 - i);
    }
#pragma unroll
    for (int j = 0; j < Ahead; ++j) {
        ahead[j] = t0 + j < t1 ? sample(t0 + j) : zero;
    }
    for (long long tb = t0; tb < t1; tb += Ring) {
#pragma unroll
        for (int jj = 0; jj < Ring; jj += 4) {
            Packed acc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                ring[jj + u]            = ahead[(jj + u) % Ahead];
                ahead[(jj + u) % Ahead] = tb + Ahead + jj + u < t1 ? sample(tb + Ahead + jj + u) : zero;
                acc[u]        = zero;
            }
#pragma unroll
            for (int q = 0; q < P; ++q) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    acc[u] = fmaV(h[q], ring[(jj + u - q + Ring) % Ring], acc[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (tb + jj + u < t1) {
                    stStream2(out + (tb + jj + u) * M + r, make_float2(packedLo(acc[u]), packedHi(acc[u])));
                }
            }
        }
    }
}

// any P: one thread per output, every tap re-reads its sample (L1/L2 absorb most of it)
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
            accRe               = __fmaf_rn(h, x.x, accRe);
            accIm               = __fmaf_rn(h, x.y, accIm);
        }
        out[o] = make_float2(accRe, accIm);
    }
}

template<int P>
void launchPfbStream(cudaStream_t s, const float2* in, const float2* state, const float* proto, float2* out, long long nFrames, int M) {
    const int       threads  = M >= 256 ? 256 : (M + 31) / 32 * 32;
    const int       gridY    = (M + threads - 1) / threads;
    static const int ctasPerSmWanted = [] { const char* e = std::getenv("GR4B200_PFB_CTAS"); return e != nullptr ? std::atoi(e) : 8; }();
    const long long  wantCtas        = static_cast<long long>(smCount()) * ctasPerSmWanted; // a few waves of the two resident CTAs per SM
    long long       stretches = wantCtas / gridY > 0 ? wantCtas / gridY : 1;
    long long       frames    = ceilDiv<long long>(nFrames, stretches);
    const long long minFrames = 16 * P; // a stretch re-reads P-1 frames: keep that below ~6 %
    frames                    = frames < minFrames ? minFrames : frames;
    frames                    = ceilDiv<long long>(frames, P + 4) * (P + 4); // whole turns of the register ring
    stretches                 = ceilDiv<long long>(nFrames, frames);
    pfbStreamKernel<P><<<dim3(static_cast<unsigned>(stretches), static_cast<unsigned>(gridY)), threads, 0, s>>>(in, state, proto, out, nFrames, M, frames);
}

// ---- fused channelizer, M = 256: polyphase FIR bank + 256-point FFT in one kernel ------------------------------------
// The filter-bank part is pfbStreamKernel's (thread = branch, register ring, look-ahead loads); instead of going to HBM
// the outputs of 16 consecutive frames are written into shared memory as 16 padded transforms (fft_radix.cuh layout),
// then the 256 threads regroup as 16 transforms x 16 threads and run the two radix-16 passes of the 256-point FFT on
// them; the spectrum goes straight to HBM. Per input sample: 8 B read + 8 B written instead of 32 B for the two stages.
// The staging area is double buffered, so one __syncthreads per 16 frames is enough.
__host__ __device__ constexpr int gcdOf(int a, int b) { return b == 0 ? a : gcdOf(b, a % b); }

template<int P>
__global__ void __launch_bounds__(256, 2) pfbChannelizer256Kernel(const float2* __restrict__ in, const float2* __restrict__ state, const float* __restrict__ proto, const float2* __restrict__ fftTables, float2* __restrict__ out, long long nFrames, long long framesPerStretch) {
    constexpr int M     = 256;
    constexpr int Ring  = P + 4;
#ifndef GR4B200_CHANNELIZER_FULL_AHEAD
#define GR4B200_CHANNELIZER_FULL_AHEAD 1
#endif
    constexpr int Ahead = GR4B200_CHANNELIZER_FULL_AHEAD ? Ring : (Ring % 8 == 0 ? 8 : Ring / 2); // look-ahead depth in frames (128 registers with a full ring turn at P = 12)
    constexpr int Turn  = Ring * 16 / gcdOf(Ring, 16);       // frames after which ring slots and FFT batches realign
    using G             = FftGeom<M>;
    static_assert(Ring % Ahead == 0 && Ring % 4 == 0, "slots must be compile-time constants");
    extern __shared__ __align__(16) unsigned char smemRaw[];
    Cx* staging0 = reinterpret_cast<Cx*>(smemRaw); // two staging areas of 16 padded transforms each

    const int            r = threadIdx.x;
    const long long      halo = static_cast<long long>(P - 1) * M;
    const long long      t0   = static_cast<long long>(blockIdx.x) * framesPerStretch;
    const long long      t1   = t0 + framesPerStretch < nFrames ? t0 + framesPerStretch : nFrames;
    const int            col  = M - 1 - r;
    const int            tr = threadIdx.x >> 4, t16 = threadIdx.x & 15; // FFT phase: transform (frame of the batch), thread in it
    float                h[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        h[q] = __ldg(proto + r + q * M);
    }
    auto sample = [&](long long t) -> Packed {
        const long long idx = t * M + col;
        const float2    v   = idx >= 0 ? ldStream2(in + idx) : __ldg(state + halo + idx);
        return packPair(v.x, v.y);
    };
    const Packed zero = packPair(0.f, 0.f);
    Packed       ring[Ring];
    Packed       ahead[Ahead];
#pragma unroll
    for (int j = 0; j < Ring; ++j) {
        ring[j] = zero;
    }
#pragma unroll
    for (int i = 0; i + 1 < P; ++i) {
        ring[Ring - 1 - i] = sample(t0 - 1 - i);
    }
#pragma unroll
    for (int j = 0; j < Ahead; ++j) {
        ahead[j] = t0 + j < t1 ? sample(t0 + j) : zero;
    }
    const int padR   = r + (r >> 4);
    int       buffer = 0; // staging area of the batch being filled; flips after every transform phase
    for (long long tb = t0; tb < t1; tb += Turn) {
#pragma unroll
        for (int f = 0; f < Turn; f += 4) {
            if (tb + (f / 16) * 16 < t1) { // the batch of 16 frames holds at least one frame of the stretch (uniform over the CTA)
                Cx*    batch = staging0 + buffer * (16 * G::kPadded);
                Packed acc[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    ring[(f + u) % Ring]   = ahead[(f + u) % Ahead];
                    ahead[(f + u) % Ahead] = tb + Ahead + f + u < t1 ? sample(tb + Ahead + f + u) : zero;
                    acc[u]                 = zero;
                }
#pragma unroll
                for (int q = 0; q < P; ++q) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        acc[u] = fmaV(h[q], ring[((f + u - q) % Ring + Ring) % Ring], acc[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    batch[((f + u) % 16) * G::kPadded + padR] = acc[u];
                }
                if ((f + 4) % 16 == 0) { // sixteen frames staged: transform them
                    __syncthreads();
                    const long long frame = tb + (f + 4 - 16) + tr;
                    Cx*             mine  = batch + tr * G::kPadded;
                    Cx              v[16];
                    fftGather<M>(t16, mine, v);
                    fftPassCompute<M, 0>(t16, v, fftTables);
                    __syncwarp();
                    fftScatter<M, 0>(t16, v, mine);
                    __syncwarp();
                    fftGather<M>(t16, mine, v);
                    fftPassCompute<M, 1>(t16, v, fftTables);
                    if (frame < t1) {
#pragma unroll
                        for (int m = 0; m < 16; ++m) {
                            float re, im;
                            cxSplit(v[m], re, im);
                            stStream2(out + frame * M + t16 + 16 * m, make_float2(re, im));
                        }
                    }
                    buffer ^= 1;
                }
            }
        }
    }
}

template<int P>
void launchChannelizer256(cudaStream_t s, const float2* in, const float2* state, const float* proto, const float2* tables, float2* out, long long nFrames) {
    constexpr int   Ring = P + 4;
    constexpr int   Turn = Ring * 16 / gcdOf(Ring, 16);
    static const int ctasPerSmWanted = [] { const char* e = std::getenv("GR4B200_PFB_CTAS"); return e != nullptr ? std::atoi(e) : 8; }();
    const long long  wantCtas  = static_cast<long long>(smCount()) * ctasPerSmWanted;
    long long       frames    = ceilDiv<long long>(nFrames, wantCtas);
    const long long minFrames = 16 * P;
    frames                    = frames < minFrames ? minFrames : frames;
    frames                    = ceilDiv<long long>(frames, Turn) * Turn;
    const long long stretches = ceilDiv<long long>(nFrames, frames);
    constexpr size_t smem = 2 * 16 * FftGeom<256>::kPadded * sizeof(Cx);
    cudaFuncSetAttribute(pfbChannelizer256Kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    pfbChannelizer256Kernel<P><<<static_cast<unsigned>(stretches), 256, smem, s>>>(in, state, proto, tables, out, nFrames, frames);
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
    int     device   = 0; // the device the plan's memory lives on
    int     M        = 0;
    int     P        = 0;
    float*  proto    = nullptr;
    float2* fftTables = nullptr; // M = 256 only: twiddles of the fused 256-point FFT (fft_radix.cuh)
    float2* state[2] = {nullptr, nullptr};
    int     current  = 0;
};

extern "C" {

gr4b200_pfb_plan* gr4b200_pfb_plan_create(const float* proto_host, size_t nChannels, size_t tapsPerBranch) {
    if (proto_host == nullptr || nChannels == 0 || tapsPerBranch == 0 || nChannels > (1u << 16) || tapsPerBranch > 4096) {
        fail("pfb_plan_create: bad arguments");
        return nullptr;
    }
    auto* plan   = new gr4b200_pfb_plan;
    plan->device = currentDevice();
    plan->M      = static_cast<int>(nChannels);
    plan->P    = static_cast<int>(tapsPerBranch);
    const size_t protoBytes = nChannels * tapsPerBranch * sizeof(float);
    const size_t haloBytes  = (tapsPerBranch - 1) * nChannels * sizeof(float2) + 16;
    bool         ok         = cudaMalloc(&plan->proto, protoBytes) == cudaSuccess && cudaMalloc(&plan->state[0], haloBytes) == cudaSuccess && cudaMalloc(&plan->state[1], haloBytes) == cudaSuccess;
    ok                      = ok && cudaMemcpy(plan->proto, proto_host, protoBytes, cudaMemcpyHostToDevice) == cudaSuccess;
    ok                      = ok && cudaMemset(plan->state[0], 0, haloBytes) == cudaSuccess && cudaMemset(plan->state[1], 0, haloBytes) == cudaSuccess;
    if (ok && nChannels == 256) {
        std::vector<float2> tables(FftGeom<256>::kTableEntries);
        fftFillTables<256>(tables.data());
        ok = cudaMalloc(&plan->fftTables, tables.size() * sizeof(float2)) == cudaSuccess && cudaMemcpy(plan->fftTables, tables.data(), tables.size() * sizeof(float2), cudaMemcpyHostToDevice) == cudaSuccess;
    }
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
    cudaFree(plan->fftTables);
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
    if (const int status = checkPlanDevice(plan->device, "pfb_filter"); status != GR4B200_OK) {
        return status;
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
    const float2* src   = reinterpret_cast<const float2*>(in);
    float2*       dst   = reinterpret_cast<float2*>(out);
    const float2* state = plan->state[plan->current];
    switch (plan->P) {
    case 4: launchPfbStream<4>(s, src, state, plan->proto, dst, static_cast<long long>(nFrames), plan->M); break;
    case 8: launchPfbStream<8>(s, src, state, plan->proto, dst, static_cast<long long>(nFrames), plan->M); break;
    case 12: launchPfbStream<12>(s, src, state, plan->proto, dst, static_cast<long long>(nFrames), plan->M); break;
    case 16: launchPfbStream<16>(s, src, state, plan->proto, dst, static_cast<long long>(nFrames), plan->M); break;
    default: pfbFilterKernel<<<static_cast<int>(want < cap ? want : cap), 256, 0, s>>>(src, state, plan->proto, dst, static_cast<long long>(nFrames), plan->M, plan->P); break;
    }
    const long long halo = static_cast<long long>(plan->P - 1) * plan->M;
    if (halo > 0) {
        pfbUpdateState<<<static_cast<int>(std::min<long long>(ceilDiv<long long>(halo, 256), cap)), 256, 0, s>>>(plan->state[plan->current], reinterpret_cast<const float2*>(in), plan->state[plan->current ^ 1], halo, total);
        plan->current ^= 1;
    }
    return checkLaunch("pfbFilterKernel", halo > 0 ? 2u : 1u);
}

int gr4b200_pfb_fused_supported(const gr4b200_pfb_plan* plan) { return plan != nullptr && plan->M == 256 && (plan->P == 4 || plan->P == 8 || plan->P == 12) ? 1 : 0; }

int gr4b200_pfb_channelizer_cf32(gr4b200_pfb_plan* plan, void* stream, const float* in, float* out, size_t nFrames) {
    if (plan == nullptr) {
        return fail("pfb_channelizer: null plan");
    }
    if (const int status = checkPlanDevice(plan->device, "pfb_channelizer"); status != GR4B200_OK) {
        return status;
    }
    if (!gr4b200_pfb_fused_supported(plan)) {
        return fail("pfb_channelizer: the fused kernel covers 256 channels with 4, 8 or 12 taps per branch; run the two stages");
    }
    if (nFrames == 0) {
        return GR4B200_OK;
    }
    if (in == nullptr || out == nullptr) {
        return fail("pfb_channelizer: null buffer");
    }
    const auto    s     = asStream(stream);
    const float2* src   = reinterpret_cast<const float2*>(in);
    float2*       dst   = reinterpret_cast<float2*>(out);
    const float2* state = plan->state[plan->current];
    switch (plan->P) {
    case 4: launchChannelizer256<4>(s, src, state, plan->proto, plan->fftTables, dst, static_cast<long long>(nFrames)); break;
    case 8: launchChannelizer256<8>(s, src, state, plan->proto, plan->fftTables, dst, static_cast<long long>(nFrames)); break;
    default: launchChannelizer256<12>(s, src, state, plan->proto, plan->fftTables, dst, static_cast<long long>(nFrames)); break;
    }
    const long long total = static_cast<long long>(nFrames) * plan->M;
    const long long halo  = static_cast<long long>(plan->P - 1) * plan->M;
    const long long cap   = static_cast<long long>(smCount()) * 8;
    pfbUpdateState<<<static_cast<int>(std::min<long long>(ceilDiv<long long>(halo, 256), cap)), 256, 0, s>>>(plan->state[plan->current], src, plan->state[plan->current ^ 1], halo, total);
    plan->current ^= 1;
    return checkLaunch("pfbChannelizer256Kernel", 2u);
}

} // extern "C"
