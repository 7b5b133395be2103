// Per-thread FIR arithmetic shared by the device kernel (fir.cu) and the host emulation used by the CPU-side tests
// (tests/host_emulation.cu). See fir.cu for the contract and the thread mapping.
#pragma once

#include <cuda_runtime.h>

#ifndef GR4B200_HD
#define GR4B200_HD __host__ __device__ __forceinline__
#endif

namespace gr4b200 {

constexpr int kLanes = 16; // 64-byte PSTL lane block / sizeof(float)

// round-to-nearest single operations that the compiler may not contract into FMAs
GR4B200_HD float fmulRn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b; // host emulation is built with -ffp-contract=off
#endif
}
GR4B200_HD float faddRn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
}
GR4B200_HD float fmaRn(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}

// ---- vector element helpers -----------------------------------------------------------------------------------------
// A complex<float> sample is handled as one packed f32x2 register pair: Blackwell's FMUL2 / FADD2 issue at the same rate
// as scalar FMUL / FADD but produce two results, so "round the product, then round the sum" (the reference's arithmetic)
// costs the same pipe time as a fused multiply-add (measured: 36.9 T mul+add pairs/s vs 18.1 T scalar, scripts/
// ubench_f32x2.cu). The real-valued stream uses the scalar forms.
using Packed = unsigned long long; // {lo = re, hi = im}

GR4B200_HD Packed packPair(float lo, float hi) { return (static_cast<Packed>(__builtin_bit_cast(unsigned, hi)) << 32) | __builtin_bit_cast(unsigned, lo); }
GR4B200_HD float  packedLo(Packed v) { return __builtin_bit_cast(float, static_cast<unsigned>(v)); }
GR4B200_HD float  packedHi(Packed v) { return __builtin_bit_cast(float, static_cast<unsigned>(v >> 32)); }

// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (observed with nvcc 12.9, also under -fmad=false), which
// would silently turn the reference's two roundings into one. The packed exact forms therefore keep at least one of the
// two operations as an explicit FMA against a RUN-TIME constant the assembler cannot fold: fma(a, b, -0) is the correctly
// rounded product (adding -0 never changes a value or a zero's sign), fma(p, 1, acc) is the correctly rounded sum; a
// plain mul.rn.f32x2 (FMUL2) next to fma(p, 1, acc), or fma(a, b, -0) next to a plain add.rn.f32x2 (FADD2), cannot fuse.
struct RoundingConsts {
    float one;     // 1.0f, read from the kernel arguments
    float negZero; // -0.0f, read from the kernel arguments
};

#ifndef GR4B200_FIR_EXACT_FORM
// 0: FFMA2(b,x,-0) + FFMA2(p,1,acc); 1: FMUL2 + FFMA2(p,1,acc); 2: FFMA2(b,x,-0) + FADD2. All three are bit-identical and
// none can be contracted; form 1 reads one register pair less per product and measured 2.8 % (full rate) / 3.1 % (/8)
// faster than form 0 on B200 (profiles/r01s_time_fir_forms.jsonl)
#define GR4B200_FIR_EXACT_FORM 1
#endif
GR4B200_HD Packed mulV(float tap, Packed x, const RoundingConsts& k) {
#ifdef __CUDA_ARCH__
    Packed t, z, d;
    asm("mov.b64 %0, {%1, %1};" : "=l"(t) : "f"(tap));
#if GR4B200_FIR_EXACT_FORM == 1
    (void)z;
    (void)k;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(t), "l"(x));
#else
    asm("mov.b64 %0, {%1, %1};" : "=l"(z) : "f"(k.negZero));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(t), "l"(x), "l"(z));
#endif
    return d;
#else
    (void)k;
    return packPair(tap * packedLo(x), tap * packedHi(x));
#endif
}
GR4B200_HD Packed addV(Packed a, Packed b, const RoundingConsts& k) {
#ifdef __CUDA_ARCH__
    Packed o, d;
#if GR4B200_FIR_EXACT_FORM == 2
    (void)o;
    (void)k;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
#else
    asm("mov.b64 %0, {%1, %1};" : "=l"(o) : "f"(k.one));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(b), "l"(o), "l"(a));
#endif
    return d;
#else
    (void)k;
    return packPair(packedLo(a) + packedLo(b), packedHi(a) + packedHi(b));
#endif
}
GR4B200_HD Packed fmaV(float tap, Packed x, Packed acc) {
#ifdef __CUDA_ARCH__
    Packed t, d;
    asm("mov.b64 %0, {%1, %1};" : "=l"(t) : "f"(tap));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(t), "l"(x), "l"(acc));
    return d;
#else
    return packPair(fmaf(tap, packedLo(x), packedLo(acc)), fmaf(tap, packedHi(x), packedHi(acc)));
#endif
}
// the same with the tap already duplicated into a register pair (t, t): the kernels that read their taps from the kernel
// parameters (TapPairs below) get them as UNIFORM register pairs -- FMUL2 / FFMA2 take a uniform operand -- so a tap costs
// neither a shared-memory access nor a vector register nor the MOV that builds the pair
GR4B200_HD Packed mulV(Packed tapPair, Packed x, const RoundingConsts& k) {
#ifdef __CUDA_ARCH__
    Packed d;
#if GR4B200_FIR_EXACT_FORM == 1
    (void)k;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(tapPair), "l"(x));
#else
    Packed z;
    asm("mov.b64 %0, {%1, %1};" : "=l"(z) : "f"(k.negZero));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(tapPair), "l"(x), "l"(z));
#endif
    return d;
#else
    return mulV(packedLo(tapPair), x, k);
#endif
}
GR4B200_HD Packed fmaV(Packed tapPair, Packed x, Packed acc) {
#ifdef __CUDA_ARCH__
    Packed d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(tapPair), "l"(x), "l"(acc));
    return d;
#else
    return fmaV(packedLo(tapPair), x, acc);
#endif
}
GR4B200_HD float mulV(Packed tapPair, float x, const RoundingConsts&) { return fmulRn(packedLo(tapPair), x); }
GR4B200_HD float fmaV(Packed tapPair, float x, float acc) { return fmaRn(packedLo(tapPair), x, acc); }
GR4B200_HD float mulV(float tap, float x, const RoundingConsts&) { return fmulRn(tap, x); }
GR4B200_HD float addV(float a, float b, const RoundingConsts&) { return faddRn(a, b); }
GR4B200_HD float fmaV(float tap, float x, float acc) { return fmaRn(tap, x, acc); }

template<typename T>
struct VecOf;
template<>
struct VecOf<float> {
    using type = float;
    static GR4B200_HD float load(const float* p) { return *p; }
    static GR4B200_HD float store(float v) { return v; }
    static GR4B200_HD float zero() { return 0.f; }
    static GR4B200_HD float negZero() { return -0.f; }
};
template<>
struct VecOf<float2> {
    using type = Packed;
    static GR4B200_HD Packed load(const float2* p) { return *reinterpret_cast<const Packed*>(p); }
    static GR4B200_HD float2 store(Packed v) { return make_float2(packedLo(v), packedHi(v)); }
    static GR4B200_HD Packed zero() { return 0ull; }
    static GR4B200_HD Packed negZero() { return 0x8000000080000000ull; }
};
GR4B200_HD float  zeroOf(float) { return 0.f; }
GR4B200_HD float2 zeroOf(float2) { return make_float2(0.f, 0.f); }

// taps of lane j are stored contiguously: tapsT[j * lanePitch + m] = b[j + 16 m]; one spare block of 8 behind the last
// row so that the "next block" tap prefetch of the last block stays in bounds
GR4B200_HD int lanePitchFor(int nTaps) { return ((nTaps + kLanes - 1) / kLanes + 7) / 8 * 8; }
// shared-memory bytes in front of the sample stages: natural-order taps + lane-major taps (+ 8 spare), rounded to 128 bytes
GR4B200_HD size_t tapsSmemBytes(int nTaps) { return (static_cast<size_t>((nTaps + 31) / 32 * 32 + kLanes * lanePitchFor(nTaps) + 8) * sizeof(float) + 127) / 128 * 128; }
// Taps in the kernel parameters (constant bank): lane-major like sTapsT, every tap stored as the pair (t, t), same pitch.
// Filters of up to kParamTaps coefficients take this route; longer ones keep the shared-memory tables.
constexpr int kParamTaps     = 256;
constexpr int kParamTapPairs = kLanes * ((kParamTaps / kLanes + 7) / 8 * 8) + 8; // 16 lanes x pitch (+ 8 spare, see lanePitchFor)
struct alignas(16) TapPairs {
    Packed pairs[kParamTapPairs];
};
inline void fillTapPairs(TapPairs& t, const float* taps, int nTaps) {
    const int pitch = lanePitchFor(nTaps);
    for (int k = 0; k < kParamTapPairs; ++k) {
        const int   j = k / pitch, m = k % pitch;
        const float v = (j < kLanes && j + kLanes * m < nTaps) ? taps[j + kLanes * m] : 0.f;
        t.pairs[k]    = packPair(v, v);
    }
}
// outputs per thread of the full-rate kernel: 16 for complex (tile = 256 threads * 16 = 4096 samples); 15 for the real
// stream, where a warp spans two segments and an odd count keeps them in different banks (tile = 3840 samples)
template<typename T>
constexpr int kOutputsPerThreadD1 = sizeof(T) == 8 ? 16 : 15;
// full-rate complex calls of at most kSmallCallTiles tiles of 4096 use tiles of 256 x kSmallCallR samples instead (fir.cu)
constexpr int kSmallCallTiles = 74, kSmallCallR = 4;
// decimating tiles (threads per CTA, outputs per thread -- odd): tile = Threads / (16/D) * 16 * R full-rate samples
constexpr int kDecimThreads2 = 256, kDecimR2 = 5;   // 2560 samples
constexpr int kDecimThreads4 = 256, kDecimR4 = 5;   // 5120
#ifndef GR4B200_DECIM8_THREADS
#define GR4B200_DECIM8_THREADS 128
#define GR4B200_DECIM8_R 5
#endif
constexpr int kDecimThreads8 = GR4B200_DECIM8_THREADS, kDecimR8 = GR4B200_DECIM8_R; // 128 x 5: 5120
// the fused DDC (mixer rotated in the staged tile, then the same window walk) prefers the smaller tile: 51 KB of shared
// memory instead of 84 KB doubles the resident CTAs and lets one CTA's rotation overlap another's convolution
// (+4 %, profiles/r01u_time_decim8_tile_variants.jsonl); the plain /8 FIR is fastest with 128 x 5
constexpr int kDecimR8Mix = 3; // 3072
constexpr int kDecimThreads16 = 128, kDecimR16 = 3; // 6144

// ---- tile layout in shared memory ---------------------------------------------------------------------------------------
// e = index into the extended tile (0 = first halo sample). Full rate: linear. Decimation D | 16: "phase major",
// element (e mod D) * pitch + e / D, because lane j of every kept output (n % D == 0) only touches samples of ONE phase
// (-j mod D): a window walk then has element stride G = 16/D inside one row, neighbouring threads read
// neighbouring elements, and with an ODD number of outputs per thread the 16/G segments of a half-warp fall into
// distinct banks (segment stride G*R elements, R odd) -- no padding, all offsets compile-time immediates.
template<typename T, int DLog2>
struct TileLayout {
    static constexpr int D = 1 << DLog2;
    int                  pitch; // elements per phase row (unused for D = 1)
    GR4B200_HD int operator()(int e) const {
        if constexpr (DLog2 == 0) {
            return e;
        } else {
            return (e & (D - 1)) * pitch + (e >> DLog2);
        }
    }
    // layout(e0 - j) - layout(e0) for an e0 that is a multiple of D (every thread's first output is): the same for all
    // threads, so a lane's pointer is "thread base + a uniform offset" and the per-lane address arithmetic runs on the
    // uniform datapath instead of in every thread
    GR4B200_HD int laneOffset(int j) const {
        if constexpr (DLog2 == 0) {
            return -j;
        } else {
            return ((-j) & (D - 1)) * pitch - ((j + D - 1) >> DLog2);
        }
    }
    // row pitch: >= cols and an odd multiple of (bank period in elements) / D, so that the staging writes of D
    // consecutive samples (one per row) by consecutive threads are conflict free as well
    static GR4B200_HD int pitchFor(int extendedTileElems) {
        if constexpr (DLog2 == 0) {
            return extendedTileElems;
        } else {
            constexpr int period = 128 / static_cast<int>(sizeof(T)); // elements per sweep over the 32 banks
            constexpr int unit   = period / D > 0 ? period / D : 1;
            const int     cols   = (extendedTileElems + D - 1) / D;
            int           mult   = (cols + unit - 1) / unit;
            mult |= 1;
            return mult * unit;
        }
    }
};

// MHere (1..8) consecutive taps of one lane against the sliding window of this thread's R outputs:
//   acc[r] (+)= b[j + 16 (mBase + m8)] * x[n0 + 16 r - j - 16 (mBase + m8)],   p[Step * q] = x[n0 - j - 16 mBase + 16 q]
// The window entry q = i - 7 serves every (r, m8) with r - m8 + 7 == i; walking i downwards visits each acc[r] in
// ascending m8, i.e. in the reference's accumulation order.
// Software pipelining (one warp alone should keep the FMA pipe busy, there are only a few warps per scheduler):
//  * window entries are fetched kAhead entries before their use; the first kAhead entries of the NEXT block (at pNext)
//    are fetched while this block finishes, so block boundaries expose no shared-memory latency: on entry w[a] holds
//    entry (R + 6 - a) of this block, on exit that of the next one;
//  * tap[] holds this block's taps; each is refilled from nextTapRow (the block that follows) right after its last use;
//  * all products of an entry are formed before they are summed.
// First: this is the lane's first block -- the m8 == 0 product STARTS the accumulator (lane[j] = f(j) + f(16 + j), no
// zero in front: pstl/unseq_backend_simd.h:468-470).
// window entries in flight per thread: few outputs per thread mean few FMAs per entry, so look further ahead
template<int R>
constexpr int kAhead = R >= 12 ? 4 : (R < 6 ? R : 6); // <= R: every block has at least R entries

template<typename T, int R, int MHere, bool Exact, bool First, int Step, typename TapE>
GR4B200_HD void firLaneBlock(const T* p, const T* pNext, TapE (&tap)[8], const TapE* nextTapRow, typename VecOf<T>::type (&w)[kAhead<R>], typename VecOf<T>::type (&acc)[R], const RoundingConsts& k) {
    using V             = VecOf<T>;
    using Vec           = typename V::type;
    constexpr int Ahead = kAhead<R>;
    constexpr int hi    = R + 6;
    constexpr int lo    = 8 - MHere;
    static_assert(hi - Ahead + 1 >= 7 && hi - lo + 1 >= Ahead, "the first entries of a block must exist for every MHere");
#pragma unroll
    for (int i = hi; i >= lo; --i) {
        const int slot = (hi - i) % Ahead;
        const Vec cur  = w[slot];
        if (i - Ahead >= lo) {
            w[slot] = V::load(p + Step * (i - Ahead - 7));
        } else { // nothing of this block left to fetch: the slot takes the next block's entry that lives there
            w[slot] = V::load(pNext + Step * (hi - slot - 7));
        }
        if constexpr (Exact) {
            Vec prod[MHere];
#pragma unroll
            for (int m8 = 0; m8 < MHere; ++m8) {
                const int r = i - 7 + m8;
                if (r >= 0 && r < R) {
                    prod[m8] = mulV(tap[m8], cur, k);
                }
            }
#pragma unroll
            for (int m8 = 0; m8 < MHere; ++m8) {
                const int r = i - 7 + m8;
                if (r >= 0 && r < R) {
                    acc[r] = (First && m8 == 0) ? prod[m8] : addV(acc[r], prod[m8], k);
                }
            }
        } else {
#pragma unroll
            for (int m8 = 0; m8 < MHere; ++m8) {
                const int r = i - 7 + m8;
                if (r >= 0 && r < R) {
                    acc[r] = fmaV(tap[m8], cur, acc[r]);
                }
            }
        }
        // tap m8 is last used by entry max(lo, 7 - m8): refill it with the next block's tap right away -- the next
        // block first needs tap m8 at its entry hi - m8. Pairs travel two at a time (one 16-byte load) once the later of
        // the two, tap m8 + 1, has had its last use.
        if constexpr (sizeof(TapE) == sizeof(Packed)) {
#pragma unroll
            for (int m8 = 0; m8 < 8; m8 += 2) {
                if (i == (6 - m8 > lo ? 6 - m8 : lo)) {
                    const ulonglong2 two = *reinterpret_cast<const ulonglong2*>(nextTapRow + m8); // rows start on 64-byte boundaries
                    tap[m8]              = two.x;
                    tap[m8 + 1]          = two.y;
                }
            }
        } else {
#pragma unroll
            for (int m8 = 0; m8 < 8; ++m8) {
                if (i == (7 - m8 > lo ? 7 - m8 : lo)) {
                    tap[m8] = nextTapRow[m8];
                }
            }
        }
    }
}

// lanes [jBegin, jEnd) all hold 8 F + L taps: F full blocks of eight and a last block of L (1..8)
template<typename T, int R, int DLog2, bool Exact, int L, typename TapE>
GR4B200_HD void firLaneGroup(int jBegin, int jEnd, int F, const T* sTile, TileLayout<T, DLog2> layout, const int* laneOffsets, int e0, const TapE* sTapsT, int pitch, TapE (&tap)[8], typename VecOf<T>::type (&w)[kAhead<R>], typename VecOf<T>::type (&total)[R], const RoundingConsts& k) {
    using V            = VecOf<T>;
    using Vec          = typename V::type;
    constexpr int Step = kLanes >> DLog2; // element distance of samples 16 apart (same phase row)
    int threadBase = layout(e0); // e0 is a multiple of D
#ifdef __CUDA_ARCH__
    asm volatile("" : "+r"(threadBase)); // keep it in a register: the compiler would otherwise rebuild it from the thread index in every lane
#endif
    const T* const pThread = sTile + threadBase;
    // the tap row and the lane-offset table are WALKED (pointer increments), not indexed by j: with `table[j]` / `j * pitch`
    // in the loop ptxas sometimes keeps j in a vector register, and every tap pair of the parameter route then arrives by
    // LDC + MOV into vector registers instead of LDCU into uniform ones (7 % on the /4 kernel, 8 % on the full-rate one)
    const int*  offsetWalk = laneOffsets != nullptr ? laneOffsets + jBegin : nullptr;
    const TapE* tapRow     = sTapsT + jBegin * pitch;
#pragma unroll 1
    for (int j = jBegin; j < jEnd; ++j, tapRow += pitch, ++offsetWalk) {
        const T*     p        = pThread + (laneOffsets != nullptr ? offsetWalk[0] : layout.laneOffset(j));
        const T*     pNextLane = pThread + (laneOffsets != nullptr ? offsetWalk[1] : layout.laneOffset(j + 1 < kLanes ? j + 1 : 0)); // entry 16 repeats entry 0
        Vec          acc[R];
        if constexpr (!Exact) { // fast: accumulate straight into the output register
#pragma unroll
            for (int r = 0; r < R; ++r) {
                acc[r] = total[r];
            }
        }
        if (F == 0) {
            firLaneBlock<T, R, L, Exact, true, Step, TapE>(p, pNextLane, tap, tapRow + pitch, w, acc, k);
        } else {
            firLaneBlock<T, R, 8, Exact, true, Step, TapE>(p, p - Step * 8, tap, tapRow + 8, w, acc, k);
#pragma unroll 1
            for (int b = 1; b < F; ++b) {
                firLaneBlock<T, R, 8, Exact, false, Step, TapE>(p - Step * 8 * b, p - Step * 8 * (b + 1), tap, tapRow + 8 * (b + 1), w, acc, k);
            }
            firLaneBlock<T, R, L, Exact, false, Step, TapE>(p - Step * 8 * F, pNextLane, tap, tapRow + pitch, w, acc, k);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            total[r] = Exact ? addV(total[r], acc[r], k) : acc[r]; // init = init + lane[j]
        }
    }
}

template<typename T, int R, int DLog2, bool Exact, typename TapE>
GR4B200_HD void firLaneGroupN(int mCount, int jBegin, int jEnd, const T* sTile, TileLayout<T, DLog2> layout, const int* laneOffsets, int e0, const TapE* sTapsT, int pitch, TapE (&tap)[8], typename VecOf<T>::type (&w)[kAhead<R>], typename VecOf<T>::type (&total)[R], const RoundingConsts& k) {
    if (jBegin >= jEnd) {
        return;
    }
    const int F = (mCount - 1) / 8;
    switch (mCount - 8 * F) { // uniform over the grid
    case 8: firLaneGroup<T, R, DLog2, Exact, 8, TapE>(jBegin, jEnd, F, sTile, layout, laneOffsets, e0, sTapsT, pitch, tap, w, total, k); break;
    case 7: firLaneGroup<T, R, DLog2, Exact, 7, TapE>(jBegin, jEnd, F, sTile, layout, laneOffsets, e0, sTapsT, pitch, tap, w, total, k); break;
    case 6: firLaneGroup<T, R, DLog2, Exact, 6, TapE>(jBegin, jEnd, F, sTile, layout, laneOffsets, e0, sTapsT, pitch, tap, w, total, k); break;
    case 5: firLaneGroup<T, R, DLog2, Exact, 5, TapE>(jBegin, jEnd, F, sTile, layout, laneOffsets, e0, sTapsT, pitch, tap, w, total, k); break;
    case 4: firLaneGroup<T, R, DLog2, Exact, 4, TapE>(jBegin, jEnd, F, sTile, layout, laneOffsets, e0, sTapsT, pitch, tap, w, total, k); break;
    case 3: firLaneGroup<T, R, DLog2, Exact, 3, TapE>(jBegin, jEnd, F, sTile, layout, laneOffsets, e0, sTapsT, pitch, tap, w, total, k); break;
    case 2: firLaneGroup<T, R, DLog2, Exact, 2, TapE>(jBegin, jEnd, F, sTile, layout, laneOffsets, e0, sTapsT, pitch, tap, w, total, k); break;
    default: firLaneGroup<T, R, DLog2, Exact, 1, TapE>(jBegin, jEnd, F, sTile, layout, laneOffsets, e0, sTapsT, pitch, tap, w, total, k); break;
    }
}

// One thread's R outputs n0 + 16 r (full-rate indices): total[r] = sum_k b[k] x[n0 + 16 r - k] in the reference order.
// sTile = the staged extended tile in `layout`; e0 = extended index of x[n0]; sTaps = the nTaps coefficients (natural
// order); sTapsT = lane-major copy.
template<typename T, int R, int DLog2, bool Exact, typename TapE = float>
GR4B200_HD void firThreadCompute(const T* sTile, TileLayout<T, DLog2> layout, int e0, const float* sTaps, const TapE* sTapsT, int nTaps, const RoundingConsts& k, T (&out)[R], const int* laneOffsets = nullptr) {
    using V            = VecOf<T>;
    using Vec          = typename V::type;
    constexpr int Step = kLanes >> DLog2;
    Vec           total[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        total[r] = V::zero(); // the reference's init = T{0}
    }
    if (nTaps > 2 * kLanes) {
        const int fullBlocks = nTaps / kLanes; // every lane has at least this many taps (>= 2)
        const int remainder  = nTaps % kLanes; // the first `remainder` lanes have one more
        const int pitch      = lanePitchFor(nTaps);
        TapE      tap[8];
        Vec       w[kAhead<R>];
        {
            if constexpr (sizeof(TapE) == sizeof(Packed)) {
#pragma unroll
                for (int m8 = 0; m8 < 8; m8 += 2) {
                    const ulonglong2 two = *reinterpret_cast<const ulonglong2*>(sTapsT + m8);
                    tap[m8]              = two.x;
                    tap[m8 + 1]          = two.y;
                }
            } else {
#pragma unroll
                for (int m8 = 0; m8 < 8; ++m8) {
                    tap[m8] = sTapsT[m8];
                }
            }
            const T* p0 = sTile + layout(e0);
#pragma unroll
            for (int a2 = 0; a2 < kAhead<R>; ++a2) {
                w[a2] = V::load(p0 + Step * (R + 6 - a2 - 7));
            }
        }
        firLaneGroupN<T, R, DLog2, Exact, TapE>(fullBlocks + 1, 0, remainder, sTile, layout, laneOffsets, e0, sTapsT, pitch, tap, w, total, k);
        firLaneGroupN<T, R, DLog2, Exact, TapE>(fullBlocks, remainder, kLanes, sTile, layout, laneOffsets, e0, sTapsT, pitch, tap, w, total, k);
    } else { // short filters: the reference folds left to right, init + f(0) + f(1) + ...
        const int shortPitch = lanePitchFor(nTaps);
        for (int tapIndex = 0; tapIndex < nTaps; ++tapIndex) {
            TapE tap;
            if constexpr (sizeof(TapE) == sizeof(float)) {
                tap = sTaps[tapIndex];
            } else { // parameter pairs are lane major: tap k sits in lane k % 16, row k / 16
                tap = sTapsT[(tapIndex % kLanes) * shortPitch + tapIndex / kLanes];
            }
            const T* p = sTile + layout(e0 - tapIndex);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const Vec w = V::load(p + Step * r);
                total[r]    = Exact ? addV(total[r], mulV(tap, w, k), k) : fmaV(tap, w, total[r]);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        out[r] = V::store(total[r]);
    }
}

// Threads: number of threads per CTA; R: outputs per thread (odd when DLog2 > 0); DLog2: log2(decimation), decimation | 16
template<typename T, int Threads, int R, int DLog2, bool Exact>
struct FirConfig {
    static_assert(DLog2 == 0 || R % 2 == 1, "decimating tiles need an odd number of outputs per thread (bank mapping)");
    static constexpr int D        = 1 << DLog2;
    static constexpr int G        = kLanes / D;           // threads per 16-sample group
    static constexpr int Segments = Threads / G;          // groups of 16*R full-rate samples per tile
    static constexpr int TileIn   = Segments * kLanes * R; // full-rate samples per tile
};

// One thread of one tile: sTile holds x[tileStart - haloPad .. tileStart + TileIn) in `layout`, outputs go to
// out[(tileStart + n)/D].
template<typename T, int Threads, int R, int DLog2, bool Exact, typename TapE = float>
GR4B200_HD void firTileThread(int tid, const T* sTile, TileLayout<T, DLog2> layout, const float* sTaps, const TapE* sTapsT, int nTaps, int haloPad, long long tileStart, long long nOut, const RoundingConsts& k, T* out, const int* laneOffsets = nullptr) {
    using Cfg      = FirConfig<T, Threads, R, DLog2, Exact>;
    const int seg  = tid / Cfg::G;
    const int tsub = tid % Cfg::G;
    const int n0   = seg * (kLanes * R) + tsub * Cfg::D; // tile-relative full-rate index of this thread's first output
    T         total[R];
    firThreadCompute<T, R, DLog2, Exact, TapE>(sTile, layout, haloPad + n0, sTaps, sTapsT, nTaps, k, total, laneOffsets);
    const long long outBase = (tileStart + n0) >> DLog2;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long o = outBase + ((kLanes * r) >> DLog2);
        if (o < nOut) {
            out[o] = total[r];
        }
    }
}

} // namespace gr4b200
