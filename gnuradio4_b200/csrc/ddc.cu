// Fused DDC: Rotator -> decimating FIR in one kernel (the reference's compile-time Merge,
// core/include/gnuradio-4.0/BlockMerging.hpp:125-138, applied as device fusion).
//
// out[j] = FIR_decim(rotator(in))[j] with the numerics of gr4b200_rotator_cf32 followed by gr4b200_fir_cf32, bit for
// bit: the decimating FIR kernel stages the RAW input tile (cp.async, phase major), rotates it in place in shared
// memory with the mixer's replayed float phases (rotator.cu checkpoints, one per 16 samples), and convolves it. The
// mixed stream exists only in shared memory: 8 B read + 8/D B written per input sample instead of 24 + 8/D.
// The FIR history carried between calls holds MIXED samples, exactly as in the unfused chain.
#include "fir_kernels.cuh"

struct gr4b200_rotator_plan;
namespace gr4b200 {
int   rotatorPrepareCheckpoints(gr4b200_rotator_plan* plan, cudaStream_t stream, size_t n, const float** runPhases); // rotator.cu
int   rotatorCommitPhase(gr4b200_rotator_plan* plan, cudaStream_t stream);
float rotatorIncrement(const gr4b200_rotator_plan* plan);
} // namespace gr4b200

using namespace gr4b200;

namespace {
// unfused fallback for decimations without a tiled kernel: the mixed samples travel through an HBM buffer owned by the
// FIR plan (same device as the plan, freed with it)
int ddcUnfused(gr4b200_rotator_plan* mixer, gr4b200_fir_plan* fir, void* stream, const float* in, float* out, size_t nIn) {
    if (nIn > fir->ddcScratchCapacity) {
        if (fir->ddcScratch != nullptr) {
            GR4B200_CUDA_TRY(cudaStreamSynchronize(asStream(stream)));
            GR4B200_CUDA_TRY(cudaFree(fir->ddcScratch));
            fir->ddcScratch         = nullptr;
            fir->ddcScratchCapacity = 0;
        }
        GR4B200_CUDA_TRY(cudaMalloc(&fir->ddcScratch, nIn * 2 * sizeof(float)));
        fir->ddcScratchCapacity = nIn;
    }
    const int status = gr4b200_rotator_cf32(mixer, stream, in, fir->ddcScratch, nIn);
    if (status != GR4B200_OK) {
        return status;
    }
    return gr4b200_fir_cf32(fir, stream, fir->ddcScratch, out, nIn);
}
} // namespace

extern "C" int gr4b200_ddc_cf32(gr4b200_rotator_plan* mixer, gr4b200_fir_plan* fir, void* stream, const float* in, float* out, size_t nIn) {
    if (mixer == nullptr || fir == nullptr) {
        return fail("ddc: null plan");
    }
    if (const int status = checkPlanDevice(fir->device, "ddc"); status != GR4B200_OK) {
        return status;
    }
    if (nIn % fir->decimate != 0) {
        return fail("ddc: nIn must be a multiple of the decimation factor", GR4B200_INSUFFICIENT_INPUT_ITEMS);
    }
    if (nIn == 0) {
        return GR4B200_OK;
    }
    if (in == nullptr || out == nullptr || reinterpret_cast<uintptr_t>(in) % 8 != 0 || reinterpret_cast<uintptr_t>(out) % 8 != 0) {
        return fail("ddc: null or misaligned buffer");
    }
    const size_t d = fir->decimate;
    if (!(d == 2 || d == 4 || d == 8 || d == 16) || fir->haloPad == 0) {
        return ddcUnfused(mixer, fir, stream, in, out, nIn);
    }
    const auto   s         = asStream(stream);
    const float* runPhases = nullptr;
    int          status    = rotatorPrepareCheckpoints(mixer, s, nIn, &runPhases);
    if (status != GR4B200_OK) {
        return status;
    }
    FirArgs args{};
    args.in        = in;
    args.out       = out;
    // the plan keeps histPad >= haloPad past samples (gr4b200_fir_plan_set_taps); the fused kernel reads and writes the last
    // haloPad of them -- the older part is not maintained by this path
    args.state     = static_cast<float2*>(fir->state[fir->current]) + (fir->histPad - fir->haloPad);
    args.newState  = static_cast<float2*>(fir->state[fir->current ^ 1]) + (fir->histPad - fir->haloPad);
    args.taps      = fir->taps;
    args.nTaps     = fir->nTaps;
    args.haloPad   = fir->haloPad;
    args.nIn       = static_cast<long long>(nIn);
    args.one       = 1.0f;
    args.negZero   = -0.0f;
    args.runPhases = runPhases;
    args.dphi      = rotatorIncrement(mixer);
    args.tapPairs  = fir->paramTaps ? &fir->tapPairs : nullptr;
    status         = fir->mode == GR4B200_FIR_EXACT ? dispatchFirDecim<float2, true, true>(s, args, d) : dispatchFirDecim<float2, false, true>(s, args, d);
    if (status != GR4B200_OK) {
        return status == GR4B200_DONE ? fail("ddc: no tiled kernel for this decimation") : status;
    }
    fir->current ^= 1;
    fir->validHistory = fir->haloPad;
    return rotatorCommitPhase(mixer, s);
}
