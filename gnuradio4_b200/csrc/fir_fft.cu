// Fused FIR -> FFT block: fir_filter (time_domain_filter.hpp:44-47) followed by the FFT block
// (blocks/fourier/include/gnuradio-4.0/fourier/fft.hpp:147-250) as ONE kernel -- the reference's compile-time Merge
// (core/include/gnuradio-4.0/BlockMerging.hpp:125-138) applied as device fusion to the metric's own flowgraph.
//
// The full-rate FIR kernel works on tiles of 4096 outputs per CTA (256 threads x 16 outputs), which is exactly one
// 4096-point transform of the FFT block behind it (256 threads x 16 points). The FIR is bound by the fp32 pipe and
// leaves HBM idle; the FFT block is bound by HBM. Fused, a CTA
//   1. convolves its tile as firKernel does (bulk-copied sample stage, lane-ordered exact sums, fir_core.cuh),
//   2. writes the 4096 filtered samples into the sample stage it has just consumed, in the padded exchange layout of
//      fft_radix.cuh (the FIR's output mapping n0 + 16 r and the FFT's gather t + 256 m are both conflict free there),
//   3. runs the three radix-16 passes (window fused) ping-ponging between that stage and one extra 34 KB array,
//   4. parks the spectrum and writes magnitude / phase / Re / Im planes (fft_epilogue.cuh).
// The filtered stream never reaches HBM: 8 B read + 16 B written per sample instead of 16 + 24, and the FFT's memory
// traffic hides under the next tile's arithmetic. Results are bit-identical to the two kernels back to back (same
// per-thread code for both halves). Requirements: complex<float>, no decimation, fftSize 4096, nIn a multiple of 4096.
#include "fft_epilogue.cuh"
#include "fft_plan.cuh"
#include "fir_kernels.cuh"

namespace gr4b200 {
namespace {

constexpr int kFusedN       = 4096;
constexpr int kFusedThreads = 256;
constexpr int kFusedR       = 16;

struct FirFftArgs {
    FirArgs       fir;
    const float*  windowT; // FFT window in the per-thread layout, or nullptr
    const float2* tables;  // FFT twiddle tables (FftGeom<4096>)
    float*        signals; // [nIn / 4096][4][4096]
    unsigned      flags;
};

// elements of one sample stage: the extended FIR tile, and at least one padded exchange array
__host__ __device__ inline int fusedStageElems(int haloPad) {
    const int fir = haloPad + kFusedN;
    return fir > FftGeom<kFusedN>::kPadded ? fir : FftGeom<kFusedN>::kPadded;
}

template<bool Exact>
__global__ void __launch_bounds__(kFusedThreads, 2) firFftBlockKernel(FirFftArgs fused) {
    using T   = float2;
    using Cfg = FirConfig<T, kFusedThreads, kFusedR, 0, Exact>;
    using G   = FftGeom<kFusedN>;
    static_assert(Cfg::TileIn == kFusedN && G::kThreads == kFusedThreads, "one FIR tile is one transform");
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ uint64_t                            fullBar[2];

    const FirArgs& args       = fused.fir;
    const int      nTaps      = args.nTaps;
    const int      haloPad    = args.haloPad;
    const int      stageElems = fusedStageElems(haloPad);
    float*         sTaps      = reinterpret_cast<float*>(smemRaw);
    float*         sTapsT     = sTaps + (nTaps + 31) / 32 * 32;
    T*             sData      = reinterpret_cast<T*>(smemRaw + tapsSmemBytes(nTaps));       // two sample stages
    Cx*            second     = reinterpret_cast<Cx*>(sData + 2 * static_cast<size_t>(stageElems)); // the other exchange array

    const T* __restrict__ in    = static_cast<const T*>(args.in);
    const T* __restrict__ state = static_cast<const T*>(args.state);
    const int       tid         = threadIdx.x;
    const RoundingConsts consts{args.one, args.negZero};
    const bool      dB        = (fused.flags & GR4B200_FFT_OUTPUT_IN_DB) != 0;
    const bool      deg       = (fused.flags & GR4B200_FFT_OUTPUT_IN_DEG) != 0;
    const bool      hasWindow = fused.windowT != nullptr;

    loadTaps<kFusedThreads>(args.taps, nTaps, sTaps, sTapsT, tid);
    if (tid == 0) {
        mbarInit(&fullBar[0], 1);
        mbarInit(&fullBar[1], 1);
        fenceBarrierInit();
    }
    __syncthreads();

    auto issueBulk = [&](long long tile, int stage) { // stage <- extended input [tileStart - haloPad, tileStart + 4096)
        const long long begin = tile * kFusedN - haloPad;
        const long long end   = tile * kFusedN + kFusedN; // nIn is a multiple of the tile
        T*              dst   = sData + static_cast<size_t>(stage) * stageElems;
        // the stage was last written through the generic proxy (exchange array of the previous transform)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbarExpectTx(&fullBar[stage], static_cast<uint32_t>((end - begin) * sizeof(T)));
        if (begin < 0) {
            bulkLoad(dst, state + (haloPad + begin), static_cast<uint32_t>(-begin * sizeof(T)), &fullBar[stage]);
            bulkLoad(dst - begin, in, static_cast<uint32_t>(end * sizeof(T)), &fullBar[stage]);
        } else {
            bulkLoad(dst, in + begin, static_cast<uint32_t>((end - begin) * sizeof(T)), &fullBar[stage]);
        }
    };

    long long tile = blockIdx.x;
    if (tid == 0 && tile < args.nTiles && args.useBulk != 0) {
        issueBulk(tile, 0);
    }
    uint32_t phaseBits = 0;

    // the FIR thread's outputs: n0 + 16 r with n0 = seg * 256 + tsub (seg = tid / 16, tsub = tid % 16)
    const int n0       = (tid >> 4) * (kLanes * kFusedR) + (tid & 15);
    const int padded0  = n0 + (n0 >> 4); // pad(n0 + 16 r) = pad(n0) + 17 r

    for (int it = 0; tile < args.nTiles; ++it, tile += gridDim.x) {
        const int       stage    = it & 1;
        const long long nextTile = tile + gridDim.x;
        if (tid == 0 && nextTile < args.nTiles && args.useBulk != 0) {
            issueBulk(nextTile, stage ^ 1); // released by the __syncthreads closing the previous iteration
        }
        T*              sTile     = sData + static_cast<size_t>(stage) * stageElems;
        const long long tileStart = tile * kFusedN;
        if (args.useBulk != 0) {
            mbarWait(&fullBar[stage], (phaseBits >> stage) & 1u);
            phaseBits ^= 1u << stage;
        } else { // input not 16-byte aligned: cooperative element-wise staging
            for (int i = tid; i < haloPad + kFusedN; i += kFusedThreads) {
                const long long q = tileStart - haloPad + i;
                sTile[i]          = q < 0 ? state[haloPad + q] : in[q];
            }
            __syncthreads();
        }

        // 1. the FIR tile, exactly as firTileThread computes it
        T filtered[kFusedR];
        firThreadCompute<T, kFusedR, 0, Exact>(sTile, TileLayout<T, 0>{stageElems}, haloPad + n0, sTaps, sTapsT, nTaps, consts, filtered);
        __syncthreads(); // every window read of this stage is done: it becomes the first exchange array

        // 2. filtered samples in natural order, padded
        Cx* first = reinterpret_cast<Cx*>(sTile);
#pragma unroll
        for (int r = 0; r < kFusedR; ++r) {
            first[padded0 + 17 * r] = cxMake(filtered[r].x, filtered[r].y);
        }
        __syncthreads();

        // 3. the transform (fftRadixKernel's passes)
        Cx v[16];
        fftGather<kFusedN>(tid, first, v);
        if (hasWindow) {
            fftApplyWindow(tid, fused.windowT, v);
        }
        fftPassCompute<kFusedN, 0>(tid, v, fused.tables);
        fftScatter<kFusedN, 0>(tid, v, second);
        __syncthreads();
        fftGather<kFusedN>(tid, second, v);
        fftPassCompute<kFusedN, 1>(tid, v, fused.tables);
        fftScatter<kFusedN, 1>(tid, v, first);
        __syncthreads();
        fftGather<kFusedN>(tid, first, v);
        fftPassCompute<kFusedN, 2>(tid, v, fused.tables);

        // 4. planes
        fftPark<kFusedN>(tid, v, second);
        __syncthreads();
        float lo[4], hi[4];
        fftBlockEpilogue<kFusedN>(tid, second, fused.signals + tile * 4 * kFusedN, dB, deg, false, lo, hi, true);
        // no barrier here: the stage was last read before the parking barrier (it may be refilled right away), and
        // `second` is next written two barriers into the following iteration -- the plane stores of this tile overlap
        // the next tile's convolution
    }
}

template<bool Exact>
int launchFirFft(cudaStream_t stream, FirFftArgs fused) {
    fused.fir.nTiles  = fused.fir.nIn / kFusedN;
    const size_t smem = tapsSmemBytes(fused.fir.nTaps) + (2 * static_cast<size_t>(fusedStageElems(fused.fir.haloPad)) + FftGeom<kFusedN>::kPadded) * sizeof(float2);
    auto         kernel = firFftBlockKernel<Exact>;
    if (smem > 227 * 1024) {
        return fail("fir_fft: filter too long for the shared-memory tile");
    }
    GR4B200_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int ctasPerSm = 0;
    GR4B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, kernel, kFusedThreads, smem));
    ctasPerSm            = ctasPerSm < 1 ? 1 : ctasPerSm;
    const long long cap  = static_cast<long long>(smCount()) * ctasPerSm;
    const int       grid = static_cast<int>(fused.fir.nTiles < cap ? fused.fir.nTiles : cap);
    kernel<<<grid, kFusedThreads, smem, stream>>>(fused);
    return checkLaunch("firFftBlockKernel");
}

} // namespace
} // namespace gr4b200

using namespace gr4b200;

extern "C" {

int gr4b200_fir_fft_fused_supported(const gr4b200_fir_plan* fir, const gr4b200_fft_plan* fft, unsigned flags) {
    return fir != nullptr && fft != nullptr && fir->decimate == 1 && fir->haloPad > 0 && fft->n == kFusedN && (flags & GR4B200_FFT_UNWRAP_PHASE) == 0 ? 1 : 0;
}

int gr4b200_fir_fft_block_cf32(gr4b200_fir_plan* fir, gr4b200_fft_plan* fft, void* stream, const float* in, size_t nIn, unsigned flags, float* signals) {
    if (!gr4b200_fir_fft_fused_supported(fir, fft, flags)) {
        return fail("fir_fft_block: the fused kernel needs a full-rate FIR with at least two taps, fftSize 4096 and no phase unwrapping; run the two blocks");
    }
    if (const int status = checkPlanDevice(fir->device, "fir_fft_block"); status != GR4B200_OK) {
        return status;
    }
    if (fft->device != fir->device) {
        return fail("fir_fft_block: the FIR and the FFT plan live on different devices");
    }
    if (nIn % kFusedN != 0) {
        return fail("fir_fft_block: nIn must be a multiple of the FFT size", GR4B200_INSUFFICIENT_INPUT_ITEMS);
    }
    if (nIn == 0) {
        return GR4B200_OK;
    }
    if (in == nullptr || signals == nullptr || reinterpret_cast<uintptr_t>(in) % 8 != 0 || reinterpret_cast<uintptr_t>(signals) % 16 != 0) {
        return fail("fir_fft_block: null or misaligned buffer");
    }
    FirFftArgs fused{};
    fused.fir.in      = in;
    fused.fir.state   = static_cast<const float2*>(fir->state[fir->current]) + (fir->histPad - fir->haloPad); // the last haloPad of the histPad samples the plan keeps
    fused.fir.taps    = fir->taps;
    fused.fir.nTaps   = fir->nTaps;
    fused.fir.haloPad = fir->haloPad;
    fused.fir.nIn     = static_cast<long long>(nIn);
    fused.fir.useBulk = reinterpret_cast<uintptr_t>(in) % 16 == 0 ? 1 : 0;
    fused.fir.one     = 1.0f;
    fused.fir.negZero = -0.0f;
    fused.windowT     = fft->windowT;
    fused.tables      = fft->tables;
    fused.signals     = signals;
    fused.flags       = flags;
    const auto s      = asStream(stream);
    const int  status = fir->mode == GR4B200_FIR_EXACT ? launchFirFft<true>(s, fused) : launchFirFft<false>(s, fused);
    if (status != GR4B200_OK) {
        return status;
    }
    firUpdateState<float2><<<ceilDiv(fir->histPad, 256), 256, 0, s>>>(static_cast<const float2*>(fir->state[fir->current]), reinterpret_cast<const float2*>(in), static_cast<float2*>(fir->state[fir->current ^ 1]), fir->histPad, static_cast<long long>(nIn));
    fir->current ^= 1;
    return checkLaunch("firUpdateState");
}

} // extern "C"
