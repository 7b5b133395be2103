// FIR and decimating FIR on complex<float> / float streams (sm_100a).
//
// Replaces fir_filter<T>::processOne (blocks/filter/include/gnuradio-4.0/filter/time_domain_filter.hpp:44-47) and
// BasicFilterProto<T, Resampling<1,1,false>>::processBulk (:190-204): y[n] = sum_k b[k] x[n-k], keep n % D == 0.
//
// Arithmetic contract (EXACT mode): the reference sums through libstdc++'s __simd_transform_reduce
// (pstl/unseq_backend_simd.h:455-505): 16 float lanes, lane j accumulates taps j, j+16, j+32, ... in that order, then
// the lanes are folded left to right onto 0.0f; products and sums are rounded separately (no FMA in the reference's
// release build). For nTaps <= 32 it is a plain left fold. The kernel reproduces exactly that order with
// __fmul_rn/__fadd_rn, so results are bit-identical. FAST mode uses fused multiply-add into one accumulator.
//
// Mapping: lane j only ever pairs output n with samples n-j-16m, so a thread that owns outputs n0, n0+16, n0+32, ...
// (R of them) sees, for a fixed j, a window of R+7 samples spaced 16 apart that slides by one entry per m. The window
// lives in registers (R+7 float2), is loaded once per (j, 8 taps) from shared memory, and feeds 8*R complex MACs.
// 16 neighbouring threads own 16 neighbouring outputs => conflict-free LDS.64 and 128-byte coalesced stores.
// Sample tiles (+ halo) are staged into shared memory by 1-D bulk async copies (cp.async.bulk, "TMA 1-D") signalled
// through an mbarrier, double buffered so the next tile streams in while the current one is being convolved.
//
// Decimation D | 16 (firDecimKernel): only outputs n % D == 0 are formed (the reference computes and drops the others,
// FilterTool.hpp:244 + time_domain_filter.hpp:190-204). Lane j of a kept output touches samples of one residue class
// mod D only, so the tile is staged PHASE MAJOR (row = index mod D) by 8-byte cp.async scatters -- coalesced on the
// global side, conflict free on the shared side -- and the same window walk runs with element stride 16/D. An odd
// number of outputs per thread spreads the segments of a half-warp over all banks (TileLayout in fir_core.cuh).
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "fir_core.cuh"

namespace gr4b200 {
namespace {

// ---- mbarrier / bulk-copy PTX ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void     mbarInit(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void     mbarExpectTx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void     mbarWait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n\t"
                 ".reg .pred p;\n\t"
                 "WAIT_LOOP:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE;\n\t"
                 "bra WAIT_LOOP;\n\t"
                 "DONE:\n\t"
                 "}" ::"r"(smemAddr(bar)),
                 "r"(parity)
                 : "memory");
}
// global -> shared bulk copy, completion counted in bytes on `bar`; all of dst/src/bytes must be multiples of 16
__device__ __forceinline__ void bulkLoad(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar) { asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar)) : "memory"); }
__device__ __forceinline__ void fenceBarrierInit() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

struct FirArgs {
    const void*  in;       // nIn samples
    void*        out;      // nIn / D samples
    const void*  state;    // haloPad samples: the haloPad inputs preceding in[0] (zeros before stream start)
    const float* taps;     // nTaps floats (global)
    int          nTaps;
    int          haloPad;  // (nTaps-1) rounded up to a multiple of 16 samples
    long long    nIn;
    long long    nTiles;
    int          useBulk;  // 1: in/state 16-byte aligned => cp.async.bulk staging
    float        one;      // 1.0f and -0.0f as run-time values, see RoundingConsts
    float        negZero;
};

template<typename T, int Threads, int R, int DLog2, bool Exact>
__global__ void __launch_bounds__(Threads) firKernel(FirArgs args) {
    using Cfg = FirConfig<T, Threads, R, DLog2, Exact>;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ uint64_t                            fullBar[2];

    const int nTaps      = args.nTaps;
    const int haloPad    = args.haloPad;
    const int stageElems = haloPad + Cfg::TileIn;
    float*    sTaps      = reinterpret_cast<float*>(smemRaw);             // natural order, nTaps (padded to 32)
    const int tapsPad    = (nTaps + 31) / 32 * 32;
    const int lanePitch  = lanePitchFor(nTaps);
    float*    sTapsT     = sTaps + tapsPad;                                // lane-major: [16][lanePitch]
    T*        sData      = reinterpret_cast<T*>(smemRaw + tapsSmemBytes(nTaps)); // 128-byte aligned

    const T* __restrict__ in    = static_cast<const T*>(args.in);
    const T* __restrict__ state = static_cast<const T*>(args.state);
    T* __restrict__ out         = static_cast<T*>(args.out);
    const long long nIn         = args.nIn;
    const long long nOut        = nIn >> DLog2;
    const int       tid         = threadIdx.x;
    const RoundingConsts consts{args.one, args.negZero};

    for (int k = tid; k < tapsPad; k += Threads) {
        sTaps[k] = k < nTaps ? args.taps[k] : 0.f;
    }
    for (int k = tid; k < kLanes * lanePitch + 8; k += Threads) { // + 8: the spare block read by the last tap prefetch
        const int j = k / lanePitch, m = k % lanePitch;
        sTapsT[k]   = (j < kLanes && j + kLanes * m < nTaps) ? args.taps[j + kLanes * m] : 0.f;
    }
    if (tid == 0) {
        mbarInit(&fullBar[0], 1);
        mbarInit(&fullBar[1], 1);
        fenceBarrierInit();
    }
    __syncthreads();

    // stage <- extended input [tileStart - haloPad, tileStart + TileIn), extended input = state ++ in (index < 0 => state)
    auto issueBulk = [&](long long tile, int stage) {
        const long long begin = tile * Cfg::TileIn - haloPad; // multiple of 16 samples
        long long       end   = tile * Cfg::TileIn + Cfg::TileIn;
        end                   = end < nIn ? end : nIn;
        T*        dst         = sData + static_cast<size_t>(stage) * stageElems;
        uint32_t  bytes       = 0;
        if (begin < 0) {
            const long long stateEnd = end < 0 ? end : 0;
            bytes += static_cast<uint32_t>((stateEnd - begin) * sizeof(T));
        }
        if (end > 0) {
            const long long inBegin = begin > 0 ? begin : 0;
            bytes += static_cast<uint32_t>((end - inBegin) * sizeof(T));
        }
        mbarExpectTx(&fullBar[stage], bytes);
        if (begin < 0) {
            const long long stateEnd = end < 0 ? end : 0;
            bulkLoad(dst, state + (haloPad + begin), static_cast<uint32_t>((stateEnd - begin) * sizeof(T)), &fullBar[stage]);
        }
        if (end > 0) {
            const long long inBegin = begin > 0 ? begin : 0;
            bulkLoad(dst + (inBegin - begin), in + inBegin, static_cast<uint32_t>((end - inBegin) * sizeof(T)), &fullBar[stage]);
        }
    };
    // a tile can be bulk-staged when the whole range is 16-byte granular: full tiles always are; the last (partial)
    // tile only when nIn*sizeof(T) is a multiple of 16
    auto bulkable = [&](long long tile) { return args.useBulk != 0 && ((tile + 1) * Cfg::TileIn <= nIn || (nIn * sizeof(T)) % 16 == 0); };

    long long tile = blockIdx.x;
    if (tid == 0 && tile < args.nTiles && bulkable(tile)) {
        issueBulk(tile, 0);
    }
    uint32_t phaseBits = 0; // bit s = parity to wait for on stage s

    for (int it = 0; tile < args.nTiles; ++it, tile += gridDim.x) {
        const int       stage    = it & 1;
        const long long nextTile = tile + gridDim.x;
        if (tid == 0 && nextTile < args.nTiles && bulkable(nextTile)) {
            issueBulk(nextTile, stage ^ 1); // that stage was released by the __syncthreads closing the previous iteration
        }
        T*              sTile     = sData + static_cast<size_t>(stage) * stageElems;
        const long long tileStart = tile * Cfg::TileIn;
        if (bulkable(tile)) {
            mbarWait(&fullBar[stage], (phaseBits >> stage) & 1u);
            phaseBits ^= 1u << stage;
        } else { // misaligned buffers or ragged tail: cooperative element-wise staging, zero fill past the end
            for (int i = tid; i < stageElems; i += Threads) {
                const long long q = tileStart - haloPad + i;
                T               v = zeroOf(T{});
                if (q < 0) {
                    v = state[haloPad + q];
                } else if (q < nIn) {
                    v = in[q];
                }
                sTile[i] = v;
            }
            __syncthreads();
        }

        firTileThread<T, Threads, R, DLog2, Exact>(tid, sTile, TileLayout<T, DLog2>{stageElems}, sTaps, sTapsT, nTaps, haloPad, tileStart, nOut, consts, out);
        __syncthreads(); // everyone is done with this stage before it is refilled
    }
}

// ---- decimating tiles: cp.async scatter into the phase-major layout --------------------------------------------------
template<int Bytes>
__device__ __forceinline__ void cpAsync(void* dstSmem, const void* srcGlobal) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smemAddr(dstSmem)), "l"(srcGlobal), "n"(Bytes) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int Pending>
__device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(Pending) : "memory"); }

template<typename T, int Threads, int R, int DLog2, bool Exact>
__global__ void __launch_bounds__(Threads) firDecimKernel(FirArgs args) {
    using Cfg    = FirConfig<T, Threads, R, DLog2, Exact>;
    using Layout = TileLayout<T, DLog2>;
    static_assert(DLog2 >= 1 && Threads % Cfg::D == 0, "decimating kernel");
    extern __shared__ __align__(128) unsigned char smemRaw[];

    const int nTaps      = args.nTaps;
    const int haloPad    = args.haloPad;
    const int extended   = haloPad + Cfg::TileIn;               // samples staged per tile
    const Layout layout{Layout::pitchFor(extended)};
    const int stageElems = Cfg::D * layout.pitch;
    float*    sTaps      = reinterpret_cast<float*>(smemRaw);
    const int tapsPad    = (nTaps + 31) / 32 * 32;
    const int lanePitch  = lanePitchFor(nTaps);
    float*    sTapsT     = sTaps + tapsPad;
    T*        sData      = reinterpret_cast<T*>(smemRaw + tapsSmemBytes(nTaps));

    const T* __restrict__ in    = static_cast<const T*>(args.in);
    const T* __restrict__ state = static_cast<const T*>(args.state);
    T* __restrict__ out         = static_cast<T*>(args.out);
    const long long nIn         = args.nIn;
    const long long nOut        = nIn >> DLog2;
    const int       tid         = threadIdx.x;
    const RoundingConsts consts{args.one, args.negZero};

    for (int k = tid; k < tapsPad; k += Threads) {
        sTaps[k] = k < nTaps ? args.taps[k] : 0.f;
    }
    for (int k = tid; k < kLanes * lanePitch + 8; k += Threads) {
        const int j = k / lanePitch, m = k % lanePitch;
        sTapsT[k]   = (j < kLanes && j + kLanes * m < nTaps) ? args.taps[j + kLanes * m] : 0.f;
    }

    // thread tid stages extended samples e = tid + k * Threads: row = tid mod D is fixed, the column advances by Threads/D
    T* const  dstBase = sData + (tid & (Cfg::D - 1)) * layout.pitch + (tid >> DLog2);
    auto      stage   = [&](long long tile, int slot) {
        T*              dst   = dstBase + static_cast<size_t>(slot) * stageElems;
        const long long first = tile * Cfg::TileIn - haloPad; // full-rate index of extended sample 0
        if (first >= 0 && first + extended <= nIn) {           // interior tile: no predicates, immediate offsets
            const T* src = in + first + tid;
            int      e   = tid;
#pragma unroll 8
            for (; e + 7 * Threads < extended; e += 8 * Threads) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    cpAsync<sizeof(T)>(dst + u * (Threads >> DLog2), src + u * Threads);
                }
                dst += 8 * (Threads >> DLog2);
                src += 8 * Threads;
            }
            for (; e < extended; e += Threads) {
                cpAsync<sizeof(T)>(dst, src);
                dst += Threads >> DLog2;
                src += Threads;
            }
        } else {
            for (int e = tid; e < extended; e += Threads, dst += Threads >> DLog2) {
                const long long q = first + e;
                if (q < 0) {
                    cpAsync<sizeof(T)>(dst, state + (haloPad + q));
                } else if (q < nIn) {
                    cpAsync<sizeof(T)>(dst, in + q);
                } else {
                    *dst = zeroOf(T{});
                }
            }
        }
        cpAsyncCommit();
    };

    long long tile = blockIdx.x;
    if (tile < args.nTiles) {
        stage(tile, 0);
    }
    for (int it = 0; tile < args.nTiles; ++it, tile += gridDim.x) {
        const int       slot     = it & 1;
        const long long nextTile = tile + gridDim.x;
        if (nextTile < args.nTiles) {
            stage(nextTile, slot ^ 1); // that slot was released by the __syncthreads closing the previous iteration
            cpAsyncWait<1>();          // everything but the group just committed has landed
        } else {
            cpAsyncWait<0>();
        }
        __syncthreads(); // all threads' parts of this tile (and the taps) are visible
        firTileThread<T, Threads, R, DLog2, Exact>(tid, sData + static_cast<size_t>(slot) * stageElems, layout, sTaps, sTapsT, nTaps, haloPad, tile * Cfg::TileIn, nOut, consts, out);
        __syncthreads(); // everyone is done with this slot before it is refilled
    }
}

// any decimation (not dividing 16): one output per thread straight from global memory, reference order. Slow path.
template<typename T, bool Exact>
__global__ void __launch_bounds__(256) firGenericKernel(const T* __restrict__ in, T* __restrict__ out, const T* __restrict__ state, const float* __restrict__ taps, int nTaps, int haloPad, long long nIn, long long decim, RoundingConsts k) {
    const long long nOut = nIn / decim;
    for (long long o = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; o < nOut; o += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long n   = o * decim;
        using V             = VecOf<T>;
        auto            x   = [&](long long q) { const T v = q < 0 ? state[haloPad + q] : in[q]; return V::load(&v); };
        auto            mac = [k](typename V::type acc, float tap, typename V::type w) { return Exact ? addV(acc, mulV(tap, w, k), k) : fmaV(tap, w, acc); };
        typename V::type sum = V::zero();
        if (nTaps > 2 * kLanes) {
            const int lastBlock = kLanes * (nTaps / kLanes);
            for (int j = 0; j < kLanes; ++j) {
                typename V::type acc = mulV(taps[j], x(n - j), k);
                for (int tapIndex = j + kLanes; tapIndex < lastBlock; tapIndex += kLanes) {
                    acc = mac(acc, taps[tapIndex], x(n - tapIndex));
                }
                if (lastBlock + j < nTaps) {
                    acc = mac(acc, taps[lastBlock + j], x(n - lastBlock - j));
                }
                sum = addV(sum, acc, k);
            }
        } else {
            for (int tapIndex = 0; tapIndex < nTaps; ++tapIndex) {
                sum = mac(sum, taps[tapIndex], x(n - tapIndex));
            }
        }
        out[o] = V::store(sum);
    }
}

// newState = last haloPad samples of (oldState ++ in[0..nIn))
template<typename T>
__global__ void firUpdateState(const T* __restrict__ oldState, const T* __restrict__ in, T* __restrict__ newState, int haloPad, long long nIn) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < haloPad; i += gridDim.x * blockDim.x) {
        const long long q = nIn - haloPad + i; // index into `in`, negative => old state
        newState[i]       = q >= 0 ? in[q] : oldState[haloPad + q];
    }
}

template<typename T, int Threads, int R, int DLog2, bool Exact>
int launchFir(cudaStream_t stream, FirArgs args) {
    using Cfg          = FirConfig<T, Threads, R, DLog2, Exact>;
    args.nTiles        = ceilDiv<long long>(args.nIn, Cfg::TileIn);
    const size_t smem    = tapsSmemBytes(args.nTaps) + 2 * static_cast<size_t>(args.haloPad + Cfg::TileIn) * sizeof(T);
    if (smem > 227 * 1024) {
        return fail("fir: filter too long for the shared-memory tile (nTaps limit ~ 10k)");
    }
    auto kernel = firKernel<T, Threads, R, DLog2, Exact>;
    if (smem > 48 * 1024) {
        GR4B200_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    }
    int ctasPerSm = 0;
    GR4B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, kernel, Threads, smem));
    ctasPerSm            = ctasPerSm < 1 ? 1 : ctasPerSm;
    const long long cap  = static_cast<long long>(smCount()) * ctasPerSm; // persistent: every CTA resident, loops over tiles
    const int       grid = static_cast<int>(args.nTiles < cap ? args.nTiles : cap);
    kernel<<<grid, Threads, smem, stream>>>(args);
    return checkLaunch("firKernel");
}

template<typename T, int Threads, int R, int DLog2, bool Exact>
int launchFirDecim(cudaStream_t stream, FirArgs args) {
    using Cfg          = FirConfig<T, Threads, R, DLog2, Exact>;
    using Layout       = TileLayout<T, DLog2>;
    args.nTiles        = ceilDiv<long long>(args.nIn, Cfg::TileIn);
    const size_t smem  = tapsSmemBytes(args.nTaps) + 2 * static_cast<size_t>(Cfg::D) * Layout::pitchFor(args.haloPad + Cfg::TileIn) * sizeof(T);
    if (smem > 227 * 1024) {
        return fail("fir: filter too long for the shared-memory tile (nTaps limit ~ 5k with decimation)");
    }
    auto kernel = firDecimKernel<T, Threads, R, DLog2, Exact>;
    if (smem > 48 * 1024) {
        GR4B200_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    }
    int ctasPerSm = 0;
    GR4B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, kernel, Threads, smem));
    ctasPerSm            = ctasPerSm < 1 ? 1 : ctasPerSm;
    const long long cap  = static_cast<long long>(smCount()) * ctasPerSm;
    const int       grid = static_cast<int>(args.nTiles < cap ? args.nTiles : cap);
    kernel<<<grid, Threads, smem, stream>>>(args);
    return checkLaunch("firDecimKernel");
}

template<typename T, bool Exact>
int dispatchFir(cudaStream_t stream, FirArgs args, size_t decimate) {
    switch (decimate) {
    case 1: return launchFir<T, 256, kOutputsPerThreadD1<T>, 0, Exact>(stream, args);
    case 2: return launchFirDecim<T, kDecimThreads2, kDecimR2, 1, Exact>(stream, args);
    case 4: return launchFirDecim<T, kDecimThreads4, kDecimR4, 2, Exact>(stream, args);
    case 8: return launchFirDecim<T, kDecimThreads8, kDecimR8, 3, Exact>(stream, args);
    case 16: return launchFirDecim<T, kDecimThreads16, kDecimR16, 4, Exact>(stream, args);
    default: {
        const long long nOut = args.nIn / static_cast<long long>(decimate);
        const int       grid = static_cast<int>(std::min<long long>(ceilDiv<long long>(nOut, 256), static_cast<long long>(smCount()) * 8));
        firGenericKernel<T, Exact><<<grid, 256, 0, stream>>>(static_cast<const T*>(args.in), static_cast<T*>(args.out), static_cast<const T*>(args.state), args.taps, args.nTaps, args.haloPad, args.nIn, static_cast<long long>(decimate), RoundingConsts{args.one, args.negZero});
        return checkLaunch("firGenericKernel");
    }
    }
}

} // namespace
} // namespace gr4b200

using namespace gr4b200;

struct gr4b200_fir_plan {
    int    nTaps    = 0;
    int    haloPad  = 0;
    size_t decimate = 1;
    int    mode     = GR4B200_FIR_EXACT;
    float* taps     = nullptr; // device
    void*  state[2] = {nullptr, nullptr}; // device, haloPad * sizeof(float2) each (ping-pong)
    int    current  = 0;
    std::vector<float> tapsHost;
};

namespace {
template<typename T>
int runFir(gr4b200_fir_plan* plan, void* stream, const float* in, float* out, size_t nIn) {
    if (plan == nullptr) {
        return fail("fir: null plan");
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
    args.state   = plan->state[plan->current];
    args.taps    = plan->taps;
    args.nTaps   = plan->nTaps;
    args.haloPad = plan->haloPad;
    args.nIn     = static_cast<long long>(nIn);
    args.useBulk = reinterpret_cast<uintptr_t>(in) % 16 == 0 ? 1 : 0;
    args.one     = 1.0f;
    args.negZero = -0.0f;
    const auto s = asStream(stream);
    const int  status = plan->mode == GR4B200_FIR_EXACT ? dispatchFir<T, true>(s, args, plan->decimate) : dispatchFir<T, false>(s, args, plan->decimate);
    if (status != GR4B200_OK) {
        return status;
    }
    if (plan->haloPad > 0) {
        firUpdateState<T><<<ceilDiv(plan->haloPad, 256), 256, 0, s>>>(static_cast<const T*>(plan->state[plan->current]), reinterpret_cast<const T*>(in), static_cast<T*>(plan->state[plan->current ^ 1]), plan->haloPad, static_cast<long long>(nIn));
        plan->current ^= 1;
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
    plan->nTaps    = static_cast<int>(nTaps);
    plan->haloPad  = static_cast<int>((nTaps - 1 + 15) / 16 * 16);
    plan->decimate = decimate;
    plan->mode     = mode == GR4B200_FIR_FAST ? GR4B200_FIR_FAST : GR4B200_FIR_EXACT;
    plan->tapsHost.assign(taps_host, taps_host + nTaps);
    const size_t stateBytes = static_cast<size_t>(plan->haloPad > 0 ? plan->haloPad : 16) * sizeof(float2);
    bool         ok         = cudaMalloc(&plan->taps, nTaps * sizeof(float)) == cudaSuccess && cudaMalloc(&plan->state[0], stateBytes) == cudaSuccess && cudaMalloc(&plan->state[1], stateBytes) == cudaSuccess;
    ok                      = ok && cudaMemcpy(plan->taps, taps_host, nTaps * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
    ok                      = ok && cudaMemset(plan->state[0], 0, stateBytes) == cudaSuccess && cudaMemset(plan->state[1], 0, stateBytes) == cudaSuccess;
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
    delete plan;
    return GR4B200_OK;
}

int gr4b200_fir_plan_reset(gr4b200_fir_plan* plan, void* stream) {
    if (plan == nullptr) {
        return fail("fir_plan_reset: null plan");
    }
    const size_t stateBytes = static_cast<size_t>(plan->haloPad > 0 ? plan->haloPad : 16) * sizeof(float2);
    return checkCuda(cudaMemsetAsync(plan->state[plan->current], 0, stateBytes, asStream(stream)), "fir_plan_reset");
}

int gr4b200_fir_cf32(gr4b200_fir_plan* plan, void* stream, const float* in, float* out, size_t nIn) { return runFir<float2>(plan, stream, in, out, nIn); }
int gr4b200_fir_f32(gr4b200_fir_plan* plan, void* stream, const float* in, float* out, size_t nIn) { return runFir<float>(plan, stream, in, out, nIn); }

} // extern "C"
