// Transforms longer than one CTA's shared memory (8192 < N <= 262144): N = N1 * N2 as two passes of COLUMN transforms
// over the same data viewed as a matrix, both built from the radix passes of fft_radix.cuh. Shared by the device kernel
// (fft_large.cu) and the host emulation (tests/host_emulation.cu).
//
// Contract as for the small sizes (algorithm/include/gnuradio-4.0/algorithm/fourier/fft.hpp:113-153): unnormalised
// forward DFT, natural order in and out. With n = N2 n1 + n2 and k = k1 + N1 k2,
//   X[k1 + N1 k2] = sum_n2 W_N2^(n2 k2) * [ W_N^(n2 k1) * sum_n1 x[N2 n1 + n2] W_N1^(n1 k1) ]
// step 1: x as an N1 x N2 row-major matrix; a CTA owns 16 adjacent columns n2, runs the 16 length-N1 transforms down
//         them, multiplies by W_N^(n2 k1) (from a table of N entries computed in double) and writes the TRANSPOSED tile
//         A[n2][k1]: 16 rows of N1 values, one contiguous 16 N1-element block (re-ordered through shared memory);
// step 2: A as an N2 x N1 matrix; a CTA owns 16 adjacent columns k1, runs the length-N2 transforms down them (over n2)
//         and stores row k2 of its tile at X[N1 k2 + k1]: natural order, no separate transpose pass.
// Column tiles are read as 128-byte row segments (16 columns x 8 bytes): lanes run along the columns (tr = tid % 16
// fastest), so one warp-wide load covers two full segments; the 16 transforms of a tile sit in 16 exchange regions
// whose pitch is odd, which keeps every 8-byte shared-memory access of a half-warp on 16 different bank pairs.
// Traffic: 2 x 16 bytes per sample (plus five twiddle-table entries per thread and pass-1 tile, served from L2).
#pragma once

#include "fft_radix.cuh"

namespace gr4b200 {

constexpr int kFftLargeMax = 1 << 18;

// N2 = 2^floor(log2(N) / 2) (second step), N1 = N / N2 (first step); both in [128, 512] for N in (8192, 262144]
GR4B200_HD int fftLargeSecond(int n) {
    int log2n = 0;
    while ((1 << log2n) < n) {
        ++log2n;
    }
    return 1 << (log2n / 2);
}

template<int L>
struct FftColumnGeom {
    static constexpr int kT       = L / 16;                     // threads per column transform
    static constexpr int kThreads = 16 * kT;                    // 16 columns per CTA: L threads
    static constexpr int kRegion  = FftGeom<L>::kPadded | 1;    // odd pitch of a column's exchange region (8-byte elements)
    static constexpr int kSmem    = 16 * kRegion * 8;
    static constexpr int kPasses  = FftGeom<L>::kPasses;
    static_assert(L >= 128 && L <= 512, "column transforms cover 128, 256 and 512 points");
};

struct FftColumnArgs {
    const Cx*     in;      // batch matrices of L rows x cols columns, row-major
    const float*  inReal;  // First only: the same matrix of REAL samples (imaginary part zero) when `in` is nullptr
    Cx*           out;     // First: batch x [cols][L]; otherwise batch x [L][cols]
    const float*  window;  // First only: natural-order window of L * cols floats, or nullptr
    const float2* twiddle; // First only: W_N^j, j in [0, N), N = L * cols
    const float2* tables;  // FftGeom<L> pass tables
    int           cols;
    long long     batch;
    int           realSpectrum; // second step of a real-input transform: bins 0 and N/2 are real (fft.hpp:245-249 sets them so)
};

GR4B200_HD Cx fftColumnLoad(const Cx* p) {
#ifdef __CUDA_ARCH__
    Cx v;
    asm volatile("ld.global.nc.L1::no_allocate.b64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
#else
    return *p;
#endif
}
GR4B200_HD void fftColumnStore(Cx* p, Cx v) {
#ifdef __CUDA_ARCH__
    asm volatile("st.global.L1::no_allocate.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#else
    *p = v;
#endif
}
GR4B200_HD float fftColumnLoadFloat(const float* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// the phases of one tile; a barrier over the CTA separates consecutive phases (the host emulation runs each phase for
// every thread before the next). `tile` = index over batch x (cols / 16).
template<int L, bool First>
GR4B200_HD void fftColumnPhaseLoad(int tid, long long tile, const FftColumnArgs& a, Cx* smem, Cx (&v)[16]) {
    using G            = FftColumnGeom<L>;
    const int       tr = tid & 15, t = tid >> 4;
    const int       tiles = a.cols / 16;
    const long long big   = tile / tiles;
    const int       c     = static_cast<int>(tile % tiles) * 16 + tr;
    bool            loaded = false;
    if constexpr (First) {
        if (a.in == nullptr) {
            const float* in = a.inReal + big * L * a.cols + c;
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                v[m] = cxMake(fftColumnLoadFloat(in + static_cast<long long>(t + G::kT * m) * a.cols), 0.f);
            }
            loaded = true;
        }
    }
    if (!loaded) {
        const Cx* in = a.in + big * L * a.cols + c;
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            v[m] = fftColumnLoad(in + static_cast<long long>(t + G::kT * m) * a.cols);
        }
    }
    if constexpr (First) {
        if (a.window != nullptr) {
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                v[m] = cxScale(v[m], fftColumnLoadFloat(a.window + static_cast<long long>(t + G::kT * m) * a.cols + c));
            }
        }
    }
    fftPassCompute<L, 0>(t, v, a.tables);
    fftScatter<L, 0>(t, v, smem + tr * G::kRegion);
}

// gather + pass P; when another pass follows the caller synchronises and calls fftColumnPhaseScatter
template<int L, int P>
GR4B200_HD void fftColumnPhasePass(int tid, const FftColumnArgs& a, const Cx* smem, Cx (&v)[16]) {
    using G      = FftColumnGeom<L>;
    const int tr = tid & 15, t = tid >> 4;
    fftGather<L>(t, smem + tr * G::kRegion, v);
    fftPassCompute<L, P>(t, v, a.tables);
}
template<int L, int P>
GR4B200_HD void fftColumnPhaseScatter(int tid, const Cx (&v)[16], Cx* smem) {
    using G      = FftColumnGeom<L>;
    const int tr = tid & 15, t = tid >> 4;
    fftScatter<L, P>(t, v, smem + tr * G::kRegion);
}

// after the last pass v[m] = Y[t + T m] of column c. Second step: straight to X[(t + T m) cols + c].
template<int L>
GR4B200_HD void fftColumnStoreRows(int tid, long long tile, const FftColumnArgs& a, const Cx (&v)[16]) {
    using G            = FftColumnGeom<L>;
    const int       tr = tid & 15, t = tid >> 4;
    const int       tiles = a.cols / 16;
    const long long big   = tile / tiles;
    const int       c     = static_cast<int>(tile % tiles) * 16 + tr;
    Cx*             out   = a.out + big * L * a.cols + c;
    const bool      realBins = a.realSpectrum != 0 && c == 0 && t == 0; // bins 0 (m = 0) and N/2 = cols * L/2 (m = 8)
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        const Cx value = (m == 0 || m == 8) && realBins ? cxMake(cxRe(v[m]), 0.f) : v[m];
        fftColumnStore(out + static_cast<long long>(t + G::kT * m) * a.cols, value);
    }
}
// First step: times W_N^(c k), parked in natural order in the column's region (every thread has gathered: barrier before).
// k = t + T m, so W_N^(c k) = W_N^(c t) * (W_N^(c T))^m: five table entries per thread (base, and the step's powers 1, 2,
// 4, 8, each rounded once from double) instead of sixteen scattered loads -- the lanes of a warp hold different columns
// c, so the direct lookups W_N^(c k) hit sixteen different sectors per request and throttled the load pipe.
template<int L>
GR4B200_HD void fftColumnTwiddlePark(int tid, long long tile, const FftColumnArgs& a, Cx (&v)[16], Cx* smem) {
    using G           = FftColumnGeom<L>;
    const int      tr = tid & 15, t = tid >> 4;
    const unsigned c  = static_cast<unsigned>(tile % (a.cols / 16)) * 16u + static_cast<unsigned>(tr);
    const unsigned mask = static_cast<unsigned>(L) * static_cast<unsigned>(a.cols) - 1u;
    const unsigned step = (c * static_cast<unsigned>(G::kT)) & mask;
    const Cx       base = cxLoadTable(a.twiddle + ((c * static_cast<unsigned>(t)) & mask));
    const Cx       u1 = cxLoadTable(a.twiddle + step), u2 = cxLoadTable(a.twiddle + ((2u * step) & mask));
    const Cx       u4 = cxLoadTable(a.twiddle + ((4u * step) & mask)), u8 = cxLoadTable(a.twiddle + ((8u * step) & mask));
    cxApplyPowers16(v, u1, u2, u4, u8);
    Cx* region = smem + tr * G::kRegion;
#pragma unroll
    for (int m = 0; m < 16; ++m) {
        region[t + G::kT * m] = cxMul(v[m], base);
    }
}
// ... and written as the contiguous block A[c0 .. c0 + 16)[0 .. L). Row e starts e * kRegion elements into the tile and
// kRegion = r (mod 16): thread tid takes element (tid - e r) mod L of row e, which puts a warp's 32 reads on one aligned
// 256-byte stretch of shared memory (two wavefronts; reading element tid straddles three).
template<int L>
GR4B200_HD void fftColumnStoreTransposed(int tid, long long tile, const FftColumnArgs& a, const Cx* smem) {
    using G             = FftColumnGeom<L>;
    static_assert(G::kThreads == L);
    const int       tiles = a.cols / 16;
    const long long big   = tile / tiles;
    Cx*             out   = a.out + big * L * a.cols + static_cast<long long>(tile % tiles) * 16 * L;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int k = (tid - e * (G::kRegion % 16)) & (L - 1);
        fftColumnStore(out + e * L + k, smem[e * G::kRegion + k]);
    }
}

} // namespace gr4b200
