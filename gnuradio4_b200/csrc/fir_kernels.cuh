// FIR kernels (sm_100a): the __global__ functions and their launchers, shared by fir.cu (FIR / decimating FIR) and
// ddc.cu (mixer fused in front of the decimating FIR). The per-thread arithmetic lives in fir_core.cuh.
//
// firKernel (full rate): sample tiles (+ halo) are staged into shared memory by 1-D bulk async copies (cp.async.bulk,
// "TMA 1-D") signalled through an mbarrier, double buffered so the next tile streams in while the current one is being
// convolved; persistent CTAs, grid = SMs x resident CTAs.
//
// firDecimKernel (decimation D | 16): only outputs n % D == 0 are formed (the reference computes and drops the others,
// FilterTool.hpp:244 + time_domain_filter.hpp:190-204). Lane j of a kept output touches samples of one residue class
// mod D only, so the tile is staged PHASE MAJOR (row = index mod D) by element-wise cp.async scatters -- coalesced on
// the global side, conflict free on the shared side -- and the same window walk runs with element stride 16/D. An odd
// number of outputs per thread spreads the segments of a half-warp over all banks (TileLayout in fir_core.cuh).
// With Mix = true the staged tile is rotated in place by the mixer before it is convolved (the DDC of SURVEY 8f.1):
// the mixed samples never travel to HBM.
#pragma once

#include <cstdlib>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>

#include "async_copy.cuh"
#include "common.cuh"
#include "fir_core.cuh"
#include "rotator_core.cuh"

namespace gr4b200 {
namespace {

struct FirArgs {
    const void*  in;       // nIn samples
    void*        out;      // nIn / D samples
    const void*  state;    // haloPad samples: the haloPad FIR inputs preceding in[0] (zeros before stream start)
    const float* taps;     // nTaps floats (global)
    int          nTaps;
    int          haloPad;  // (nTaps-1) rounded up to a multiple of 16 samples
    long long    nIn;
    long long    nTiles;
    int          useBulk;  // 1: in/state 16-byte aligned => cp.async.bulk staging
    float        one;      // 1.0f and -0.0f as run-time values, see RoundingConsts
    float        negZero;
    // Mix only: the mixer in front of the filter
    const float* runPhases; // phase in front of sample 16 r of this call, r = 0 .. nIn/16 (rotator.cu checkpoints)
    float        dphi;
    void*        newState;  // haloPad samples: the last haloPad MIXED samples of state ++ mix(in), written by the kernel
    // host side only: the plan's taps as (t, t) pairs for the kernels that read them from their parameters; nullptr => shared-memory tables
    const TapPairs* tapPairs;
    // layout(e0 - j) - layout(e0), j = 0 .. 16 (entry 16 = entry 0), for the tile layout of the kernel being launched
    // (TileLayout::laneOffset): read through the uniform datapath instead of being computed in every thread
    int laneOffsets[kLanes + 1];
};
template<typename T, int DLog2>
inline void fillLaneOffsets(FirArgs& args, int extendedTileElems) {
    const TileLayout<T, DLog2> layout{TileLayout<T, DLog2>::pitchFor(extendedTileElems)};
    for (int j = 0; j <= kLanes; ++j) {
        args.laneOffsets[j] = layout.laneOffset(j % kLanes);
    }
}

// Where a kernel reads its taps: scalar floats from shared memory (every product then builds its (t, t) pair with a MOV),
// ready-made (t, t) pairs from shared memory (16-byte loads, two taps each), or pairs from the kernel parameters
// (uniform registers: neither shared-memory traffic nor vector registers).
constexpr int kTapsSmemScalar = 0, kTapsSmemPairs = 1, kTapsParamPairs = 2;
// second kernel parameter: the tap pairs (kTapsParamPairs) or nothing
struct NoTapPairs {};
template<int TapMode>
using TapParam = std::conditional_t<TapMode == kTapsParamPairs, TapPairs, NoTapPairs>;
template<int TapMode>
using TapElem = std::conditional_t<TapMode == kTapsSmemScalar, float, Packed>;
// shared-memory bytes of the tap tables in front of the sample stages
template<int TapMode>
GR4B200_HD size_t tapTableBytes(int nTaps) {
    if constexpr (TapMode == kTapsParamPairs) {
        return 0;
    } else if constexpr (TapMode == kTapsSmemPairs) { // natural-order floats (short filters) + lane-major pairs (+ 8 spare)
        return (static_cast<size_t>((nTaps + 31) / 32 * 32) * sizeof(float) + static_cast<size_t>(kLanes * lanePitchFor(nTaps) + 8) * sizeof(Packed) + 127) / 128 * 128;
    } else {
        return tapsSmemBytes(nTaps);
    }
}

// tap tables at the front of dynamic shared memory: natural order (padded to 32) + lane major (+ 8 spare, see lanePitchFor)
template<int Threads>
__device__ __forceinline__ void loadTaps(const float* __restrict__ taps, int nTaps, float* sTaps, float* sTapsT, int tid) {
    const int tapsPad   = (nTaps + 31) / 32 * 32;
    const int lanePitch = lanePitchFor(nTaps);
    for (int k = tid; k < tapsPad; k += Threads) {
        sTaps[k] = k < nTaps ? taps[k] : 0.f;
    }
    for (int k = tid; k < kLanes * lanePitch + 8; k += Threads) {
        const int j = k / lanePitch, m = k % lanePitch;
        sTapsT[k]   = (j < kLanes && j + kLanes * m < nTaps) ? taps[j + kLanes * m] : 0.f;
    }
}

template<int Threads>
__device__ __forceinline__ void loadTapPairs(const float* __restrict__ taps, int nTaps, float* sTaps, Packed* sPairs, int tid) {
    const int tapsPad   = (nTaps + 31) / 32 * 32;
    const int lanePitch = lanePitchFor(nTaps);
    for (int k = tid; k < tapsPad; k += Threads) {
        sTaps[k] = k < nTaps ? taps[k] : 0.f;
    }
    for (int k = tid; k < kLanes * lanePitch + 8; k += Threads) {
        const int   j = k / lanePitch, m = k % lanePitch;
        const float v = (j < kLanes && j + kLanes * m < nTaps) ? taps[j + kLanes * m] : 0.f;
        sPairs[k]     = packPair(v, v);
    }
}

// ---- full rate -------------------------------------------------------------------------------------------------------
template<typename T, int Threads, int R, bool Exact, int TapMode>
__global__ void __launch_bounds__(Threads, 2) firKernel(const __grid_constant__ FirArgs args, const __grid_constant__ TapParam<TapMode> tapParam) {
    using Cfg = FirConfig<T, Threads, R, 0, Exact>;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    __shared__ uint64_t                            fullBar[2];

    const int nTaps      = args.nTaps;
    const int haloPad    = args.haloPad;
    const int stageElems = haloPad + Cfg::TileIn;
    float*    sTaps      = reinterpret_cast<float*>(smemRaw);
    float*    sTapsT     = sTaps + (nTaps + 31) / 32 * 32;
    T*        sData      = reinterpret_cast<T*>(smemRaw + tapTableBytes<TapMode>(nTaps)); // 128-byte aligned
    const TapElem<TapMode>* tapTable;
    if constexpr (TapMode == kTapsParamPairs) {
        tapTable = tapParam.pairs;
    } else if constexpr (TapMode == kTapsSmemPairs) {
        tapTable = reinterpret_cast<const Packed*>(sTapsT);
    } else {
        tapTable = sTapsT;
    }

    const T* __restrict__ in    = static_cast<const T*>(args.in);
    const T* __restrict__ state = static_cast<const T*>(args.state);
    T* __restrict__ out         = static_cast<T*>(args.out);
    const long long nIn         = args.nIn;
    const int       tid         = threadIdx.x;
    const RoundingConsts consts{args.one, args.negZero};

    gridDependencyLaunch(); // the next kernel of the stream may be set up while this one runs (common.cuh)
    if constexpr (TapMode == kTapsSmemScalar) {
        loadTaps<Threads>(args.taps, nTaps, sTaps, sTapsT, tid);
    } else if constexpr (TapMode == kTapsSmemPairs) {
        loadTapPairs<Threads>(args.taps, nTaps, sTaps, reinterpret_cast<Packed*>(sTapsT), tid);
    }
    if (tid == 0) {
        mbarInit(&fullBar[0], 1);
        mbarInit(&fullBar[1], 1);
        fenceBarrierInit();
    }
    __syncthreads();
    gridDependencyWait(); // the samples (and the space the outputs go to) belong to the previous kernel until here

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

        firTileThread<T, Threads, R, 0, Exact, TapElem<TapMode>>(tid, sTile, TileLayout<T, 0>{stageElems}, sTaps, tapTable, nTaps, haloPad, tileStart, nIn, consts, out); // full rate: the lane offset is just -j (and a table lookup in the lane loop makes ptxas keep j, and with it the tap pairs, out of the uniform registers)
        __syncthreads(); // everyone is done with this stage before it is refilled
    }
}

// ---- decimating tiles: cp.async scatter into the phase-major layout (+ the mixer in place) ---------------------------

// Rotator<std::complex<float>>::processOne (blocks/math/.../Rotator.hpp:51-61) applied in place to the staged tile.
// A group = 8 consecutive extended samples: for D = 8 that is one column of the phase-major tile (row = sample within
// the group), so neighbouring threads touch neighbouring elements. The phase in front of a group comes from the
// checkpoint of its 16-sample run (the second half of a run replays the first 8 steps without using them). A thread
// advances all of its groups together: the phase recurrence is serial, independent groups hide its latency.
// Samples in front of the call (q < 0: the carried FIR history) are already mixed; samples past the end do not exist.
// checkpoints of this thread's groups, fetched before the tile itself is waited for
template<int Threads, int PerThread>
__device__ __forceinline__ void mixLoadCheckpoints(float (&phase)[PerThread], int groupBase, int groups, long long first, long long nIn, const float* __restrict__ runPhases, int tid) {
#pragma unroll
    for (int b = 0; b < PerThread; ++b) {
        const int       g  = groupBase + tid + b * Threads;
        const long long q0 = first + 8 * g; // multiple of 8 (`first` is a multiple of 16): a group never straddles q = 0
        static_assert(kRun == 8, "one checkpoint per group of 8 samples");
        phase[b]           = (g < groups && q0 >= 0 && q0 < nIn) ? __ldg(runPhases + (q0 >> 3)) : 0.f;
    }
}

// General form: every sample is checked (inside the call? phase in the fast sin/cos range? NaN recovery of the product).
template<int Threads, int DLog2, int PerThread>
__device__ __noinline__ void mixTileChecked(float2* sTile, TileLayout<float2, DLog2> layout, const float* __restrict__ runPhases, int groupBase, int groups, long long first, long long nIn, float dphi, int tid) {
#pragma unroll 1
    for (int b = 0; b < PerThread; ++b) {
        const int       g  = groupBase + tid + b * Threads;
        const long long q0 = first + 8 * g;
        if (g >= groups || q0 < 0 || q0 >= nIn) {
            continue;
        }
        float ph = __ldg(runPhases + (q0 >> 3)); // fetched here (rare path) so that the callers' checkpoint registers never have to live in memory
        for (int i = 0; i < 8 && q0 + i < nIn; ++i) {
            bool wrapped;
            ph = stepPhase(ph, dphi, wrapped); // Rotator.hpp:52-58: increment first, then use
            float sn, cs;
            mixerSinCos(ph, &sn, &cs);
            float2*      p = sTile + layout(8 * g + i);
            const float2 x = *p;
            *p             = complexMulAnnexG(x.x, x.y, cs, sn);
        }
    }
}

// Interior tiles (every sample inside the call): straight-line code over all of the thread's groups. Returns false when a
// sample needs the general form (a checkpoint phase outside [0, 2 pi_f], or a product with a NaN part); the caller then
// re-stages the thread's raw samples and takes the checked path -- rare (non-finite data, a start phase out of range).
// Positive = sign of dphi: inside [0, 2 pi_f] only one of the two wrap tests can fire (stepPhaseInRange), and the phase
// is never -0 (mixerSinCosInRange).
template<int Threads, int DLog2, int PerThread, bool Positive>
__device__ __forceinline__ bool mixTileFast(float2* sTile, TileLayout<float2, DLog2> layout, const float (&phaseIn)[PerThread], int groupBase, int groups, float dphi, int tid) {
    float phase[PerThread];
    bool  ok = true;
#pragma unroll
    for (int b = 0; b < PerThread; ++b) {
        phase[b] = phaseIn[b];
        ok       = ok && phase[b] >= 0.f && phase[b] <= kTwoPi;
    }
    ok = ok && fabsf(dphi) <= 3.1415927f && dphi != 0.f;
    // slots 0 .. PerThread-2 always hold a group (the callers size PerThread that way); only the last slot can lie past
    // the end of the tile
    const bool lastLive = groupBase + tid + (PerThread - 1) * Threads < groups;
    // element of sample i of slot b's group; for D = 8 a group is one column of the phase-major tile (sample i sits i rows down)
    auto at = [&](int b, int i) -> float2& {
        const int g = groupBase + tid + b * Threads;
        if constexpr (DLog2 == 3) {
            return sTile[i * layout.pitch + g];
        } else {
            return sTile[layout(8 * g + i)];
        }
    };
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        float2 x[PerThread];
#pragma unroll
        for (int b = 0; b < PerThread; ++b) {
            x[b] = (b < PerThread - 1 || lastLive) ? at(b, i) : make_float2(0.f, 0.f);
        }
        float sn[PerThread], cs[PerThread];
#pragma unroll
        for (int b = 0; b < PerThread; ++b) {
            phase[b] = stepPhaseInRange<Positive>(phase[b], dphi);
        }
#pragma unroll
        for (int b = 0; b < PerThread; ++b) { // FP64 pipe: leaves the fp32 pipe to the tap products
            mixerSinCosInRange(phase[b], &sn[b], &cs[b]);
        }
#pragma unroll
        for (int b = 0; b < PerThread; ++b) {
            const float ac = __fmul_rn(x[b].x, cs[b]), bd = __fmul_rn(x[b].y, sn[b]), ad = __fmul_rn(x[b].x, sn[b]), bc = __fmul_rn(x[b].y, cs[b]);
            const float re = __fsub_rn(ac, bd), im = __fadd_rn(ad, bc);
            ok             = ok && !(re != re || im != im); // either part NaN (a superset of Annex G's "both"): general form
            x[b]           = make_float2(re, im);
        }
#pragma unroll
        for (int b = 0; b < PerThread; ++b) {
            if (b < PerThread - 1 || lastLive) {
                at(b, i) = x[b];
            }
        }
    }
    return ok;
}

// The mixer over groups [groupFrom, groups) of a staged tile; phase[] holds the checkpoints of the first pass (fetched by the
// caller before it waited for the tile). interior: every one of those samples lies inside the call.
template<int Threads, int DLog2, int PerThread>
__device__ __forceinline__ void mixGroups(float2* sTile, TileLayout<float2, DLog2> layout, float (&phase)[PerThread], int groupFrom, int groups, long long first, long long nIn, bool interior, const float2* __restrict__ in, const float* __restrict__ runPhases, float dphi, int tid) {
    for (int groupBase = groupFrom; groupBase < groups; groupBase += PerThread * Threads) { // one pass unless the filter is very long
        if (groupBase > groupFrom) {
            mixLoadCheckpoints<Threads, PerThread>(phase, groupBase, groups, first, nIn, runPhases, tid);
        }
        if (interior && groupBase == groupFrom) {
            const bool ok = dphi > 0.f ? mixTileFast<Threads, DLog2, PerThread, true>(sTile, layout, phase, groupBase, groups, dphi, tid) : mixTileFast<Threads, DLog2, PerThread, false>(sTile, layout, phase, groupBase, groups, dphi, tid);
            if (!ok) {
                for (int b = 0; b < PerThread; ++b) { // this thread's groups again, from the raw input
                    const int g = groupBase + tid + b * Threads;
                    for (int i = 0; i < 8 && g < groups; ++i) {
                        sTile[layout(8 * g + i)] = in[first + 8 * g + i];
                    }
                }
                mixTileChecked<Threads, DLog2, PerThread>(sTile, layout, runPhases, groupBase, groups, first, nIn, dphi, tid);
            }
        } else {
            mixTileChecked<Threads, DLog2, PerThread>(sTile, layout, runPhases, groupBase, groups, first, nIn, dphi, tid);
        }
    }
}

// Stages = 2: the next tile streams in while the current one is convolved (few, large CTAs). Stages = 1: a CTA stages,
// waits and convolves in turn and the overlap comes from the OTHER CTAs resident on the SM -- half the shared memory per
// CTA, so twice the resident warps for the same tile shape (the decimating kernels are bound by warps per scheduler:
// fixed-latency `wait` stalls at 2 warps per scheduler, profiles/r01z_fir_decim8_ncu.md).
template<typename T, int Threads, int R, int DLog2, bool Exact, bool Mix, int Stages, int TapMode>
__global__ void __launch_bounds__(Threads) firDecimKernel(const __grid_constant__ FirArgs args, const __grid_constant__ TapParam<TapMode> tapParam) {
    static_assert(Stages == 1 || Stages == 2, "single or double buffered tile");
    constexpr bool Prefetch = Stages == 2;
    using Cfg    = FirConfig<T, Threads, R, DLog2, Exact>;
    using Layout = TileLayout<T, DLog2>;
    static_assert(DLog2 >= 1 && Threads % Cfg::D == 0, "decimating kernel");
    static_assert(!Mix || sizeof(T) == 8, "the mixer works on complex samples");
    extern __shared__ __align__(128) unsigned char smemRaw[];

    const int    nTaps      = args.nTaps;
    const int    haloPad    = args.haloPad;
    const int    extended   = haloPad + Cfg::TileIn; // samples staged per tile
    const Layout layout{Layout::pitchFor(extended)};
    const int    stageElems = Cfg::D * layout.pitch;
    float*       sTaps      = reinterpret_cast<float*>(smemRaw);
    float*       sTapsT     = sTaps + (nTaps + 31) / 32 * 32;
    T*           sData      = reinterpret_cast<T*>(smemRaw + tapTableBytes<TapMode>(nTaps));
    const TapElem<TapMode>* tapTable;
    if constexpr (TapMode == kTapsParamPairs) {
        tapTable = tapParam.pairs;
    } else if constexpr (TapMode == kTapsSmemPairs) {
        tapTable = reinterpret_cast<const Packed*>(sTapsT);
    } else {
        tapTable = sTapsT;
    }

    const T* __restrict__ in    = static_cast<const T*>(args.in);
    const T* __restrict__ state = static_cast<const T*>(args.state);
    T* __restrict__ out         = static_cast<T*>(args.out);
    const long long nIn         = args.nIn;
    const long long nOut        = nIn >> DLog2;
    const int       tid         = threadIdx.x;
    const RoundingConsts consts{args.one, args.negZero};

    gridDependencyLaunch();
    if constexpr (TapMode == kTapsSmemScalar) {
        loadTaps<Threads>(args.taps, nTaps, sTaps, sTapsT, tid);
    } else if constexpr (TapMode == kTapsSmemPairs) {
        loadTapPairs<Threads>(args.taps, nTaps, sTaps, reinterpret_cast<Packed*>(sTapsT), tid);
    }
    gridDependencyWait(); // samples, checkpoints and output space belong to the previous kernel until here

    // thread tid stages extended samples e = tid + k * Threads: row = tid mod D is fixed, the column advances by Threads/D
    T* const dstBase = sData + (tid & (Cfg::D - 1)) * layout.pitch + (tid >> DLog2);
    // eStart (a multiple of 16): first extended sample to fetch -- 0, or haloPad when the halo was carried over in shared memory
    auto     stage   = [&](long long tile, int slot, int eStart = 0) {
        T*              dst   = dstBase + static_cast<size_t>(slot) * stageElems + (eStart >> DLog2);
        const long long first = tile * Cfg::TileIn - haloPad; // full-rate index of extended sample 0
        if (first + eStart >= 0 && first + extended <= nIn) {  // interior tile: no predicates, immediate offsets
            const T* src = in + first + eStart + tid;
            int      e   = eStart + tid;
#pragma unroll 1
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
            for (int e = eStart + tid; e < extended; e += Threads, dst += Threads >> DLog2) {
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

    // mixer: groups of 8 samples per thread (the halo pad depends on the filter, so the count is a run-time bound)
    constexpr int PerThread = Mix ? (Cfg::TileIn / 8 + 16 + Threads - 1) / Threads : 1; // one pass for halos <= 128 samples
    const int     groups    = extended / 8;
    // carry for the next call: the last haloPad mixed samples of state ++ mix(in). Every tile writes the part it holds
    // (tiles overlap by the halo, the values agree); only the last tile(s) hold any of it.
    auto writeNewState = [&](const T* sTile, long long first) {
        if (first + extended > nIn - haloPad) {
            T* __restrict__ newState = static_cast<T*>(args.newState);
            for (int e = tid; e < extended; e += Threads) {
                const long long q = first + e;
                if (q >= nIn - haloPad && q < nIn) {
                    newState[q - (nIn - haloPad)] = sTile[layout(e)];
                }
            }
        }
    };

    if constexpr (Mix && Stages == 1) {
        // Every CTA owns a CONTIGUOUS range of tiles: the halo of tile k+1 is the mixed tail of tile k, already in shared
        // memory -- it is moved to the front (one element per thread) instead of being fetched and rotated again, and the
        // mixer's groups then divide evenly over the threads (TileIn/8 per tile instead of TileIn/8 + haloPad/8).
        constexpr int   PerThreadNew = (Cfg::TileIn / 8 + Threads - 1) / Threads;
        const long long tilesPerCta  = (args.nTiles + gridDim.x - 1) / gridDim.x;
        const long long tileBegin    = static_cast<long long>(blockIdx.x) * tilesPerCta;
        const long long tileEnd      = tileBegin + tilesPerCta < args.nTiles ? tileBegin + tilesPerCta : args.nTiles;
        constexpr int   CarryPer     = Threads >= 128 ? 1 : 128 / Threads; // halo elements a thread moves (halos of up to 128 samples)
        const bool      canCarry     = haloPad <= CarryPer * Threads && haloPad <= Cfg::TileIn;
        T* const        sTile        = sData;
        T               carried[CarryPer];
        for (long long tile = tileBegin; tile < tileEnd; ++tile) {
            const long long first    = tile * Cfg::TileIn - haloPad;
            const bool      interior = first >= 0 && first + extended <= nIn;
            if (canCarry && tile > tileBegin) {
#pragma unroll
                for (int c = 0; c < CarryPer; ++c) {
                    if (tid + c * Threads < haloPad) {
                        sTile[layout(tid + c * Threads)] = carried[c]; // read before the barrier that closed the previous tile
                    }
                }
                stage(tile, 0, haloPad);
                float phase[PerThreadNew];
                mixLoadCheckpoints<Threads, PerThreadNew>(phase, haloPad / 8, groups, first, nIn, args.runPhases, tid);
                cpAsyncWait<0>();
                __syncthreads();
                mixGroups<Threads, DLog2, PerThreadNew>(sTile, layout, phase, haloPad / 8, groups, first, nIn, interior, in, args.runPhases, args.dphi, tid);
            } else {
                stage(tile, 0);
                float phase[PerThread];
                mixLoadCheckpoints<Threads, PerThread>(phase, 0, groups, first, nIn, args.runPhases, tid);
                cpAsyncWait<0>();
                __syncthreads();
                mixGroups<Threads, DLog2, PerThread>(sTile, layout, phase, 0, groups, first, nIn, interior, in, args.runPhases, args.dphi, tid);
            }
            __syncthreads();
            writeNewState(sTile, first);
            firTileThread<T, Threads, R, DLog2, Exact, TapElem<TapMode>>(tid, sTile, layout, sTaps, tapTable, nTaps, haloPad, tile * Cfg::TileIn, nOut, consts, out, args.laneOffsets);
#pragma unroll
            for (int c = 0; c < CarryPer; ++c) {
                if (canCarry && tid + c * Threads < haloPad) {
                    carried[c] = sTile[layout(Cfg::TileIn + tid + c * Threads)]; // nobody writes the tile while it is being convolved
                }
            }
            __syncthreads(); // everyone is done with the tile before it is refilled
        }
        return;
    }

    long long tile = blockIdx.x;
    if (Prefetch && tile < args.nTiles) {
        stage(tile, 0);
    }
    for (int it = 0; tile < args.nTiles; ++it, tile += gridDim.x) {
        const int       slot     = Prefetch ? (it & 1) : 0;
        const long long nextTile = tile + gridDim.x;
        const long long first    = tile * Cfg::TileIn - haloPad;
        float           phase[PerThread];
        if constexpr (Prefetch) {
            if (nextTile < args.nTiles) {
                stage(nextTile, slot ^ 1); // that slot was released by the __syncthreads closing the previous iteration
            }
        } else {
            stage(tile, 0);
        }
        if constexpr (Mix) {
            mixLoadCheckpoints<Threads, PerThread>(phase, 0, groups, first, nIn, args.runPhases, tid);
        }
        if (Prefetch && nextTile < args.nTiles) {
            cpAsyncWait<1>(); // everything but the group just committed has landed
        } else {
            cpAsyncWait<0>();
        }
        __syncthreads(); // all threads' parts of this tile (and the taps) are visible
        T* sTile = sData + static_cast<size_t>(slot) * stageElems;
        if constexpr (Mix) {
            mixGroups<Threads, DLog2, PerThread>(sTile, layout, phase, 0, groups, first, nIn, first >= 0 && first + extended <= nIn, in, args.runPhases, args.dphi, tid);
            __syncthreads();
            writeNewState(sTile, first);
        }
        firTileThread<T, Threads, R, DLog2, Exact, TapElem<TapMode>>(tid, sTile, layout, sTaps, tapTable, nTaps, haloPad, tile * Cfg::TileIn, nOut, consts, out, args.laneOffsets);
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

// Every CTA loops over tiles (the next tile is prefetched while the current one is computed). Grid = `defaultMult` x the
// resident grid: CTAs that live for the whole launch finish unevenly and leave SMs idle at the end; sixteen waves of
// shorter loops let the hardware scheduler even that out -- exact FIR 63.1 -> 65.3 GS/s, fast 101.5 -> 107.1, /8 exact
// 338 -> 351, fused DDC 170 -> 183; only the /8 fast kernel is best with the resident grid (profiles/r01z_time_fir_grid_mult.jsonl).
template<typename Kernel, typename TapArg>
int launchPersistent(Kernel kernel, const char* name, cudaStream_t stream, const FirArgs& args, const TapArg& tapArg, int threads, size_t smem, int defaultMult) {
    if (smem > 227 * 1024) {
        return fail("fir: filter too long for the shared-memory tile (nTaps limit: a few thousand)");
    }
    // the shared-memory opt-in and the occupancy query are driver calls: once per (kernel, device, shared-memory size),
    // not once per work chunk -- a streaming flowgraph launches this thousands of times per second. The opt-in is a
    // property of the kernel, not of the launch: it is only ever raised (a filter with fewer taps must not lower it under
    // a plan that is still in use).
    static std::mutex                                          cacheMutex;
    static std::map<std::tuple<const void*, int, size_t>, int> occupancyCache;
    static std::map<std::pair<const void*, int>, size_t>       optedIn;
    const void* const kernelKey = reinterpret_cast<const void*>(kernel);
    const int         device    = currentDevice();
    int               ctasPerSm = 0;
    bool              raise     = false;
    {
        std::lock_guard<std::mutex> lock(cacheMutex);
        if (const auto it = occupancyCache.find(std::make_tuple(kernelKey, device, smem)); it != occupancyCache.end()) {
            ctasPerSm = it->second;
        }
        size_t& current = optedIn[std::make_pair(kernelKey, device)];
        if (smem > 48 * 1024 && smem > current) {
            current = smem;
            raise   = true;
        }
    }
    if (raise) {
        GR4B200_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    }
    if (ctasPerSm == 0) {
        GR4B200_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSm, kernel, threads, smem));
        ctasPerSm = ctasPerSm < 1 ? 1 : ctasPerSm;
        std::lock_guard<std::mutex> lock(cacheMutex);
        occupancyCache[std::make_tuple(kernelKey, device, smem)] = ctasPerSm;
    }
    // GR4B200_FIR_GRID_MULT: grid in units of the resident grid (A/B timing of shorter-lived CTAs; 0 = one CTA per tile)
    static const int envMult  = [] { const char* e = std::getenv("GR4B200_FIR_GRID_MULT"); return e != nullptr ? std::atoi(e) : -1; }();
    const int        gridMult = envMult >= 0 ? envMult : defaultMult;
    const long long  cap      = gridMult > 0 ? static_cast<long long>(smCount()) * ctasPerSm * gridMult : args.nTiles;
    const int        grid     = static_cast<int>(args.nTiles < cap ? args.nTiles : cap);
    const cudaError_t launched = launchDependent(kernel, dim3(static_cast<unsigned>(grid)), dim3(static_cast<unsigned>(threads)), smem, stream, args, tapArg);
    if (launched != cudaSuccess) {
        return checkCuda(launched, name);
    }
    return checkLaunch(name);
}

// GR4B200_FIR_PARAM_TAPS=0: keep the shared-memory tap tables (A/B timing; results are the same bits)
// GR4B200_FIR_TAP_MODE=0|1|2 overrides the tap source of the complex kernels (A/B timing; results are the same bits)
template<typename T>
inline int tapModeFor(const FirArgs& args, int preferred) {
    if constexpr (sizeof(T) != 8) {
        return kTapsSmemScalar; // the real-valued stream keeps the scalar tables (nothing to pair)
    }
    static const int forced = [] { const char* e = std::getenv("GR4B200_FIR_TAP_MODE"); return e != nullptr ? std::atoi(e) : -1; }();
    const int        mode   = forced >= 0 ? forced : preferred;
    return mode == kTapsParamPairs && args.tapPairs == nullptr ? kTapsSmemPairs : mode;
}

template<typename T, int Threads, int R, bool Exact>
int launchFir(cudaStream_t stream, FirArgs args) {
    using Cfg         = FirConfig<T, Threads, R, 0, Exact>;
    args.nTiles       = ceilDiv<long long>(args.nIn, Cfg::TileIn);
    // GR4B200_FIR_EXTRA_SMEM=bytes: occupancy experiment (more shared memory per CTA = fewer resident CTAs); results do not change
    static const size_t extra = [] { const char* e = std::getenv("GR4B200_FIR_EXTRA_SMEM"); return e != nullptr ? static_cast<size_t>(std::atol(e)) : size_t{0}; }();
    const size_t data = 2 * static_cast<size_t>(args.haloPad + Cfg::TileIn) * sizeof(T) + extra;
    if constexpr (sizeof(T) == 8) {
        // measured (profiles/r02m_time_variants.jsonl, r02q): the exact kernel gains from parameter taps (63.2 -> 66.5 GS/s), the
        // fast one (half the arithmetic per tap, so twice the uniform loads per FMA) loses (106 -> 94; shared pairs: 102)
        const int mode = tapModeFor<T>(args, Exact ? kTapsParamPairs : kTapsSmemScalar);
        if (mode == kTapsParamPairs) {
            return launchPersistent(firKernel<T, Threads, R, Exact, kTapsParamPairs>, "firKernel", stream, args, *args.tapPairs, Threads, data, 16);
        }
        if (mode == kTapsSmemPairs) {
            return launchPersistent(firKernel<T, Threads, R, Exact, kTapsSmemPairs>, "firKernel", stream, args, NoTapPairs{}, Threads, tapTableBytes<kTapsSmemPairs>(args.nTaps) + data, 16);
        }
    }
    return launchPersistent(firKernel<T, Threads, R, Exact, kTapsSmemScalar>, "firKernel", stream, args, NoTapPairs{}, Threads, tapsSmemBytes(args.nTaps) + data, 16);
}

template<typename T, int Threads, int R, int DLog2, bool Exact, bool Mix, int Stages = 2>
int launchFirDecim(cudaStream_t stream, FirArgs args) {
    using Cfg         = FirConfig<T, Threads, R, DLog2, Exact>;
    using Layout      = TileLayout<T, DLog2>;
    args.nTiles       = ceilDiv<long long>(args.nIn, Cfg::TileIn);
    fillLaneOffsets<T, DLog2>(args, args.haloPad + Cfg::TileIn);
    const size_t data = Stages * static_cast<size_t>(Cfg::D) * Layout::pitchFor(args.haloPad + Cfg::TileIn) * sizeof(T);
    // sixteen waves of CTAs (shorter-lived CTAs even out the tail; with the fused mixer every CTA owns a contiguous range of
    // tiles, still several per CTA in long calls: 224 GS/s at two waves, 232 at sixteen, profiles/r02q_time_variants.jsonl)
    const int    mult = Exact || Mix || Stages == 1 ? 16 : 1;
    if constexpr (sizeof(T) == 8) {
        // warp-sized CTAs: parameter taps (no tap table per CTA, more CTAs per SM; /8 exact 412 GS/s against 372 with scalar
        // shared taps and 359 with shared pairs); CTA-wide tiles: scalar shared taps measured best (/2 112 vs 108 with shared
        // pairs vs 104 with parameter taps, /4 212 / 207 / 211, /16 398 / 388 / 376). The fused DDC takes shared PAIRS: with the
        // mixer in the same kernel ptxas loads parameter taps with LDC into vector registers (+ MOVs) instead of LDCU into
        // uniform ones -- 210 GS/s against 230 with shared pairs and 225 with scalars (profiles/r02z_time_tap_modes.jsonl)
        const int mode = tapModeFor<T>(args, Threads == 32 ? (Mix ? kTapsSmemPairs : kTapsParamPairs) : kTapsSmemScalar);
        if (mode == kTapsParamPairs) {
            return launchPersistent(firDecimKernel<T, Threads, R, DLog2, Exact, Mix, Stages, kTapsParamPairs>, "firDecimKernel", stream, args, *args.tapPairs, Threads, data, mult);
        }
        if (mode == kTapsSmemPairs) {
            return launchPersistent(firDecimKernel<T, Threads, R, DLog2, Exact, Mix, Stages, kTapsSmemPairs>, "firDecimKernel", stream, args, NoTapPairs{}, Threads, tapTableBytes<kTapsSmemPairs>(args.nTaps) + data, mult);
        }
    }
    return launchPersistent(firDecimKernel<T, Threads, R, DLog2, Exact, Mix, Stages, kTapsSmemScalar>, "firDecimKernel", stream, args, NoTapPairs{}, Threads, tapsSmemBytes(args.nTaps) + data, mult);
}

// decimation D | 16 with the tile shapes of fir_core.cuh; returns GR4B200_DONE (never a valid launch status here) when
// `decimate` has no tiled kernel
template<typename T, bool Exact, bool Mix>
int dispatchFirDecim(cudaStream_t stream, const FirArgs& args, size_t decimate) {
    // the one-warp tiles fetch their halo once per warp: fine while the halo is a fraction of the tile (127 taps: 128 of
    // 960 .. 1536 samples), wasteful for long filters -- those keep the CTA-wide tiles, whose halo is shared by 4-8 warps
    const bool shortHalo = args.haloPad <= 256;
    switch (decimate) {
    case 2: {
        if constexpr (sizeof(T) == 8 && Exact && !Mix) { // GR4B200_DECIM2_VARIANT: tile experiments
            static const int variantEnv = [] { const char* e = std::getenv("GR4B200_DECIM2_VARIANT"); return e != nullptr ? std::atoi(e) : -1; }();
            const int        variant    = shortHalo ? variantEnv : 0;
            // one warp per CTA, single stage (see /8 below): 121 -> 129 GS/s (profiles/r02s_time_variants.jsonl)
            switch (variant) {
            case 0: return launchFirDecim<T, kDecimThreads2, kDecimR2, 1, Exact, Mix>(stream, args);
            default: return launchFirDecim<T, 32, 15, 1, Exact, Mix, 1>(stream, args);
            }
        }
        return launchFirDecim<T, kDecimThreads2, kDecimR2, 1, Exact, Mix>(stream, args);
    }
    case 4: {
        if constexpr (sizeof(T) == 8 && Exact && !Mix) { // GR4B200_DECIM4_VARIANT: tile experiments
            static const int variantEnv = [] { const char* e = std::getenv("GR4B200_DECIM4_VARIANT"); return e != nullptr ? std::atoi(e) : -1; }();
            const int        variant    = shortHalo ? variantEnv : 0;
            // one warp per CTA, single stage (see /8 below): 210 -> 249 GS/s (profiles/r02r_time_variants.jsonl)
            switch (variant) {
            case 0: return launchFirDecim<T, kDecimThreads4, kDecimR4, 2, Exact, Mix>(stream, args);
            default: return launchFirDecim<T, 32, 9, 2, Exact, Mix, 1>(stream, args);
            }
        }
        return launchFirDecim<T, kDecimThreads4, kDecimR4, 2, Exact, Mix>(stream, args);
    }
    case 8: {
        // GR4B200_DECIM8_VARIANT: tile shape experiments (threads x outputs per thread x stages), complex kernels only.
        // Measured (profiles/r02p_time_variants.jsonl, 2^28 samples): exact /8 352 GS/s with CTA-wide double-buffered tiles
        // (128 x 5 x 2), 369 with 128 x 7 x 1, 407 with ONE WARP PER CTA (32 x 5 x 1, taps from the parameters): a warp stages,
        // waits for and convolves its own 16 segments, every barrier is a warp barrier, ~18 such pipelines run per SM
        // independently of one another; the halo is fetched once per warp (+10 % staging). Fused DDC: 192 -> 213.
        if constexpr (sizeof(T) == 8) {
            static const int variantEnv = [] { const char* e = std::getenv("GR4B200_DECIM8_VARIANT"); return e != nullptr ? std::atoi(e) : -1; }();
            const int        variant    = shortHalo ? variantEnv : (Mix ? 2 : 0);
            switch (variant) {
            case 0: return launchFirDecim<T, 128, 5, 3, Exact, Mix, 2>(stream, args);
            case 2: return launchFirDecim<T, 128, 5, 3, Exact, Mix, 1>(stream, args);
            case 4: return launchFirDecim<T, 128, 7, 3, Exact, Mix, 1>(stream, args);
            case 13: return launchFirDecim<T, 32, 7, 3, Exact, Mix, 1>(stream, args);
            case 16: return launchFirDecim<T, 64, 5, 3, Exact, Mix, 1>(stream, args);
            default: return launchFirDecim<T, 32, 5, 3, Exact, Mix, 1>(stream, args);
            }
        } else {
            return launchFirDecim<T, kDecimThreads8, kDecimR8, 3, Exact, Mix, 2>(stream, args);
        }
    }
    case 16: {
        if constexpr (sizeof(T) == 8 && Exact && !Mix) { // GR4B200_DECIM16_VARIANT: tile experiments
            static const int variantEnv = [] { const char* e = std::getenv("GR4B200_DECIM16_VARIANT"); return e != nullptr ? std::atoi(e) : -1; }();
            const int        variant    = shortHalo ? variantEnv : 0;
            // one warp per CTA, single stage: 391 -> 538 GS/s (profiles/r02s_time_variants.jsonl)
            switch (variant) {
            case 0: return launchFirDecim<T, kDecimThreads16, kDecimR16, 4, Exact, Mix>(stream, args);
            case 2: return launchFirDecim<T, 32, 5, 4, Exact, Mix, 1>(stream, args);
            default: return launchFirDecim<T, 32, 3, 4, Exact, Mix, 1>(stream, args);
            }
        }
        return launchFirDecim<T, kDecimThreads16, kDecimR16, 4, Exact, Mix>(stream, args);
    }
    default: return GR4B200_DONE;
    }
}

} // namespace
} // namespace gr4b200

// the plan behind gr4b200_fir_plan_create (fir.cu), also read by the fused DDC (ddc.cu)
struct gr4b200_fir_plan {
    int    device   = 0;                  // the device the plan's memory lives on
    int    nTaps    = 0;
    int    haloPad  = 0;
    size_t decimate = 1;
    int    mode     = GR4B200_FIR_EXACT;
    gr4b200::TapPairs tapPairs{};         // host: (t, t) pairs, lane major, passed to the kernels as a parameter when paramTaps
    bool   paramTaps = false;             // nTaps <= kParamTaps
    float* taps     = nullptr;            // device
    void*  state[2] = {nullptr, nullptr}; // device, histPad * sizeof(float2) each (ping-pong); the kernels see the last haloPad samples
    int    refCapacity  = 32;             // the reference's HistoryBuffer capacity: 32, or bit_ceil(nTaps) once a longer filter was set
    int    histPad      = 0;              // samples kept = max(haloPad, (refCapacity - 1) rounded up to 16): a later, longer `b` that fits finds its past
    int    validHistory = 0;              // how many of them are true history (the fused DDC maintains only haloPad)
    size_t tapsCapacity = 0;              // floats allocated behind `taps`
    int    current  = 0;
    float2* olsSpectrum       = nullptr;  // overlap-save mode: FFT_4096(taps) / 4096 in the kernel's per-thread layout
    float2* olsTables         = nullptr;  //                    twiddle tables of the 4096-point passes
    float* ddcScratch         = nullptr;  // unfused DDC fallback (ddc.cu): mixed samples of one call
    size_t ddcScratchCapacity = 0;        // in samples
};
