// FIR and decimating FIR on complex<float> / float streams (sm_100a): plans and C ABI. Kernels: fir_kernels.cuh.
//
// Replaces fir_filter<T>::processOne (blocks/filter/include/gnuradio-4.0/filter/time_domain_filter.hpp:44-47) and
// BasicFilterProto<T, Resampling<1,1,false>>::processBulk (:190-204): y[n] = sum_k b[k] x[n-k], keep n % D == 0.
//
// Arithmetic contract (EXACT mode): the reference sums through libstdc++'s __simd_transform_reduce
// (pstl/unseq_backend_simd.h:455-505): 16 float lanes, lane j accumulates taps j, j+16, j+32, ... in that order (the
// first product starts the lane), then the lanes are folded left to right onto 0.0f; products and sums are rounded
// separately (no FMA in the reference's release build). For nTaps <= 32 it is a plain left fold. The kernels reproduce
// exactly that order, so results are bit-identical. FAST mode uses fused multiply-add into one accumulator.
//
// Mapping: lane j only ever pairs output n with samples n-j-16m, so a thread that owns outputs n0, n0+16, n0+32, ...
// (R of them) sees, for a fixed j, a window of R+7 samples spaced 16 apart that slides by one entry per m. Each window
// entry is loaded once per (j, 8 taps) from shared memory and feeds up to 8 complex MACs.
#include <vector>

#include "fir_kernels.cuh"
#include "fir_overlap_save.cuh"

using namespace gr4b200;

namespace {
size_t bitCeil(size_t v) { // std::bit_ceil (C++20) for this C++17 translation unit
    size_t p = 1;
    while (p < v) {
        p <<= 1;
    }
    return p;
}
template<typename T, bool Exact>
int dispatchFir(cudaStream_t stream, const FirArgs& args, size_t decimate) {
    if (decimate == 1) {
        // short calls (a streaming work chunk of 64 Ki samples is 16 tiles of 4096: 16 of 148 SMs busy for the ~12 us one tile
        // takes): quarter-size tiles spread the same chunk over four times as many SMs -- the call is bound by one tile's
        // latency, not by the pipe
        if constexpr (sizeof(T) == 8) {
            if (args.nIn <= kSmallCallTiles * 4096LL) {
                return launchFir<T, 256, kSmallCallR, Exact>(stream, args);
            }
        }
        return launchFir<T, 256, kOutputsPerThreadD1<T>, Exact>(stream, args);
    }
    const int status = dispatchFirDecim<T, Exact, false>(stream, args, decimate);
    if (status != GR4B200_DONE) {
        return status;
    }
    const long long nOut = args.nIn / static_cast<long long>(decimate);
    const int       grid = static_cast<int>(std::min<long long>(ceilDiv<long long>(nOut, 256), static_cast<long long>(smCount()) * 8));
    firGenericKernel<T, Exact><<<grid, 256, 0, stream>>>(static_cast<const T*>(args.in), static_cast<T*>(args.out), static_cast<const T*>(args.state), args.taps, args.nTaps, args.haloPad, args.nIn, static_cast<long long>(decimate), RoundingConsts{args.one, args.negZero});
    return checkLaunch("firGenericKernel");
}
} // namespace

namespace {
template<typename T>
int runFir(gr4b200_fir_plan* plan, void* stream, const float* in, float* out, size_t nIn, bool historyInStream = false) {
    if (plan == nullptr) {
        return fail("fir: null plan");
    }
    if (const int status = checkPlanDevice(plan->device, "fir"); status != GR4B200_OK) {
        return status;
    }
    if (nIn % plan->decimate != 0) { // the reference fixes input_chunk_size = decimate (time_domain_filter.hpp:166-168)
        return fail("fir: nIn must be a multiple of the decimation factor", GR4B200_INSUFFICIENT_INPUT_ITEMS);
    }
    if (nIn == 0) {
        return GR4B200_OK;
    }
    if (in == nullptr || out == nullptr || reinterpret_cast<uintptr_t>(in) % sizeof(T) != 0 || reinterpret_cast<uintptr_t>(out) % sizeof(T) != 0) {
        return fail("fir: null or misaligned buffer");
    }
    FirArgs args{};
    args.in      = in;
    args.out     = out;
    // the plan keeps histPad >= haloPad past samples (see gr4b200_fir_plan_set_taps); the kernels read the last haloPad of them
    args.state   = historyInStream ? static_cast<const void*>(reinterpret_cast<const T*>(in) - plan->haloPad) : static_cast<const void*>(static_cast<const T*>(plan->state[plan->current]) + (plan->histPad - plan->haloPad));
    args.taps    = plan->taps;
    args.nTaps   = plan->nTaps;
    args.haloPad = plan->haloPad;
    args.nIn     = static_cast<long long>(nIn);
    args.useBulk = reinterpret_cast<uintptr_t>(in) % 16 == 0 ? 1 : 0;
    args.one      = 1.0f;
    args.negZero  = -0.0f;
    args.tapPairs = plan->paramTaps ? &plan->tapPairs : nullptr;
    const auto s  = asStream(stream);
    int        status;
    if (plan->mode == GR4B200_FIR_OVERLAP_SAVE) {
        if constexpr (std::is_same_v<T, float2>) {
            OlsArgs ols{};
            ols.in       = reinterpret_cast<const float2*>(in);
            ols.state    = static_cast<const float2*>(args.state);
            ols.out      = reinterpret_cast<float2*>(out);
            ols.spectrum = plan->olsSpectrum;
            ols.tables   = plan->olsTables;
            ols.nIn      = static_cast<long long>(nIn);
            ols.overlap  = (plan->nTaps - 1 + 1) / 2 * 2; // even: every window then starts on a 16-byte boundary
            ols.hop      = kOlsN - ols.overlap;
            ols.useBulk  = reinterpret_cast<uintptr_t>(in) % 16 == 0 ? 1 : 0;
            ols.haloPad  = plan->haloPad;
            status       = launchOverlapSave(s, ols);
        } else {
            return fail("fir: the overlap-save mode is implemented for complex<float> streams");
        }
    } else {
        status = plan->mode == GR4B200_FIR_EXACT ? dispatchFir<T, true>(s, args, plan->decimate) : dispatchFir<T, false>(s, args, plan->decimate);
    }
    if (status != GR4B200_OK) {
        return status;
    }
    if (plan->histPad > 0 && !historyInStream) {
        firUpdateState<T><<<ceilDiv(plan->histPad, 256), 256, 0, s>>>(static_cast<const T*>(plan->state[plan->current]), reinterpret_cast<const T*>(in), static_cast<T*>(plan->state[plan->current ^ 1]), plan->histPad, static_cast<long long>(nIn));
        plan->current ^= 1;
        plan->validHistory = plan->histPad;
        return checkLaunch("firUpdateState");
    }
    return GR4B200_OK;
}
} // namespace

extern "C" {

gr4b200_fir_plan* gr4b200_fir_plan_create(const float* taps_host, size_t nTaps, size_t decimate, int mode) {
    if (taps_host == nullptr || nTaps == 0 || nTaps > (1u << 20) || decimate == 0) {
        fail("fir_plan_create: need taps, 1 <= nTaps <= 2^20, decimate >= 1");
        return nullptr;
    }
    auto* plan     = new gr4b200_fir_plan;
    plan->device   = currentDevice();
    plan->nTaps    = static_cast<int>(nTaps);
    plan->haloPad  = static_cast<int>((nTaps - 1 + 15) / 16 * 16);
    plan->decimate = decimate;
    plan->mode     = mode == GR4B200_FIR_FAST ? GR4B200_FIR_FAST : (mode == GR4B200_FIR_OVERLAP_SAVE ? GR4B200_FIR_OVERLAP_SAVE : GR4B200_FIR_EXACT);
    if (plan->mode == GR4B200_FIR_OVERLAP_SAVE && (decimate != 1 || nTaps > static_cast<size_t>(kOlsMaxTaps))) {
        fail("fir_plan_create: the overlap-save mode needs decimate == 1 and nTaps <= 2049");
        delete plan;
        return nullptr;
    }
    plan->paramTaps = nTaps <= static_cast<size_t>(kParamTaps);
    if (plan->paramTaps) {
        fillTapPairs(plan->tapPairs, taps_host, plan->nTaps);
    }
    plan->refCapacity  = nTaps > 32 ? static_cast<int>(bitCeil(nTaps)) : 32;
    plan->histPad      = std::max(plan->haloPad, (plan->refCapacity - 1 + 15) / 16 * 16);
    plan->validHistory = plan->histPad; // zeros: what the reference's fresh HistoryBuffer holds
    plan->tapsCapacity = nTaps;
    const size_t stateBytes = static_cast<size_t>(plan->histPad) * sizeof(float2);
    bool         ok         = cudaMalloc(&plan->taps, nTaps * sizeof(float)) == cudaSuccess && cudaMalloc(&plan->state[0], stateBytes) == cudaSuccess && cudaMalloc(&plan->state[1], stateBytes) == cudaSuccess;
    ok                      = ok && cudaMemcpy(plan->taps, taps_host, nTaps * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
    ok                      = ok && cudaMemset(plan->state[0], 0, stateBytes) == cudaSuccess && cudaMemset(plan->state[1], 0, stateBytes) == cudaSuccess;
    if (ok && plan->mode == GR4B200_FIR_OVERLAP_SAVE) {
        std::vector<float2> spectrum, tables(FftGeom<kOlsN>::kTableEntries);
        olsSpectrum(taps_host, nTaps, spectrum);
        fftFillTables<kOlsN>(tables.data());
        ok = cudaMalloc(&plan->olsSpectrum, spectrum.size() * sizeof(float2)) == cudaSuccess && cudaMemcpy(plan->olsSpectrum, spectrum.data(), spectrum.size() * sizeof(float2), cudaMemcpyHostToDevice) == cudaSuccess;
        ok = ok && cudaMalloc(&plan->olsTables, tables.size() * sizeof(float2)) == cudaSuccess && cudaMemcpy(plan->olsTables, tables.data(), tables.size() * sizeof(float2), cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (!ok) {
        checkCuda(cudaGetLastError(), "fir_plan_create");
        gr4b200_fir_plan_destroy(plan);
        return nullptr;
    }
    return plan;
}

int gr4b200_fir_plan_destroy(gr4b200_fir_plan* plan) {
    if (plan == nullptr) {
        return GR4B200_OK;
    }
    cudaFree(plan->taps);
    cudaFree(plan->state[0]);
    cudaFree(plan->state[1]);
    cudaFree(plan->ddcScratch);
    cudaFree(plan->olsSpectrum);
    cudaFree(plan->olsTables);
    delete plan;
    return GR4B200_OK;
}

int gr4b200_fir_plan_reset(gr4b200_fir_plan* plan, void* stream) {
    if (plan == nullptr) {
        return fail("fir_plan_reset: null plan");
    }
    const size_t stateBytes = static_cast<size_t>(plan->histPad) * sizeof(float2);
    plan->validHistory      = plan->histPad;
    return checkCuda(cudaMemsetAsync(plan->state[plan->current], 0, stateBytes, asStream(stream)), "fir_plan_reset");
}

int gr4b200_fir_plan_set_taps(gr4b200_fir_plan* plan, void* stream, const float* taps_host, size_t nTaps) {
    if (plan == nullptr || taps_host == nullptr || nTaps == 0 || nTaps > (1u << 20)) {
        return fail("fir_plan_set_taps: need a plan and 1 <= nTaps <= 2^20 coefficients");
    }
    if (const int status = checkPlanDevice(plan->device, "fir_plan_set_taps"); status != GR4B200_OK) {
        return status;
    }
    if (plan->mode == GR4B200_FIR_OVERLAP_SAVE) {
        return fail("fir_plan_set_taps: the overlap-save mode precomputes the filter's spectrum; create a new plan");
    }
    const auto s = asStream(stream);
    GR4B200_CUDA_TRY(cudaStreamSynchronize(s)); // launches that still read the old coefficients finish first (a settings change is rare)
    const int newHalo = static_cast<int>((nTaps - 1 + 15) / 16 * 16);
    if (nTaps > static_cast<size_t>(plan->refCapacity)) {
        // does not fit the reference's buffer: a new one, bit_ceil(nTaps) long, filled with zeros (time_domain_filter.hpp:40-42)
        plan->refCapacity   = static_cast<int>(bitCeil(nTaps));
        const int newPad    = std::max(newHalo, (plan->refCapacity - 1 + 15) / 16 * 16);
        const size_t bytes  = static_cast<size_t>(newPad) * sizeof(float2);
        void*        fresh[2] = {nullptr, nullptr};
        if (cudaMalloc(&fresh[0], bytes) != cudaSuccess || cudaMalloc(&fresh[1], bytes) != cudaSuccess) {
            cudaFree(fresh[0]);
            return checkCuda(cudaGetLastError(), "fir_plan_set_taps");
        }
        GR4B200_CUDA_TRY(cudaMemset(fresh[0], 0, bytes));
        GR4B200_CUDA_TRY(cudaMemset(fresh[1], 0, bytes));
        cudaFree(plan->state[0]);
        cudaFree(plan->state[1]);
        plan->state[0] = fresh[0], plan->state[1] = fresh[1];
        plan->current      = 0;
        plan->histPad      = newPad;
        plan->validHistory = newPad;
    } else if (newHalo > plan->validHistory) {
        // fits, but the samples beyond validHistory were not maintained (fused DDC calls): they read as zeros
        // (only complex streams get here: the DDC is the one path that leaves validHistory below histPad)
        GR4B200_CUDA_TRY(cudaMemset(plan->state[plan->current], 0, static_cast<size_t>(plan->histPad - plan->validHistory) * sizeof(float2)));
    }
    if (nTaps > plan->tapsCapacity) {
        float* fresh = nullptr;
        GR4B200_CUDA_TRY(cudaMalloc(&fresh, nTaps * sizeof(float)));
        cudaFree(plan->taps);
        plan->taps         = fresh;
        plan->tapsCapacity = nTaps;
    }
    GR4B200_CUDA_TRY(cudaMemcpy(plan->taps, taps_host, nTaps * sizeof(float), cudaMemcpyHostToDevice));
    plan->nTaps     = static_cast<int>(nTaps);
    plan->haloPad   = newHalo;
    plan->paramTaps = nTaps <= static_cast<size_t>(kParamTaps);
    if (plan->paramTaps) {
        fillTapPairs(plan->tapPairs, taps_host, plan->nTaps);
    }
    return GR4B200_OK;
}

int gr4b200_fir_cf32(gr4b200_fir_plan* plan, void* stream, const float* in, float* out, size_t nIn) { return runFir<float2>(plan, stream, in, out, nIn); }
int gr4b200_fir_f32(gr4b200_fir_plan* plan, void* stream, const float* in, float* out, size_t nIn) { return runFir<float>(plan, stream, in, out, nIn); }
size_t gr4b200_fir_plan_history_items(const gr4b200_fir_plan* plan) { return plan == nullptr ? 0 : static_cast<size_t>(plan->haloPad); }
int gr4b200_fir_cf32_contiguous(gr4b200_fir_plan* plan, void* stream, const float* in, float* out, size_t nIn) { return runFir<float2>(plan, stream, in, out, nIn, true); }
int gr4b200_fir_f32_contiguous(gr4b200_fir_plan* plan, void* stream, const float* in, float* out, size_t nIn) { return runFir<float>(plan, stream, in, out, nIn, true); }

} // extern "C"
