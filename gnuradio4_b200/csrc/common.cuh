// Shared helpers for the gr4b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/gr4b200.h"

namespace gr4b200 {

// thread-local last-error string behind gr4b200_last_error()
void        setLastError(const std::string& message);
const char* lastError();

inline int fail(const std::string& message, int status = GR4B200_ERROR) {
    setLastError(message);
    return status;
}

inline int checkCuda(cudaError_t err, const char* what) {
    if (err == cudaSuccess) {
        return GR4B200_OK;
    }
    return fail(std::string(what) + ": " + cudaGetErrorString(err));
}

#define GR4B200_CUDA_TRY(expr)                                       \
    do {                                                             \
        const int _status = ::gr4b200::checkCuda((expr), #expr);     \
        if (_status != GR4B200_OK) {                                 \
            return _status;                                          \
        }                                                            \
    } while (0)

// launch-error check: every kernel launch in the C ABI ends with this (cudaGetLastError -> work::Status::ERROR)
inline int checkLaunch(const char* kernelName) { return checkCuda(cudaGetLastError(), kernelName); }

inline cudaStream_t asStream(void* stream) { return static_cast<cudaStream_t>(stream); }

int smCount(); // SMs of the current device (cached per device)

template<typename T>
constexpr T ceilDiv(T a, T b) {
    return (a + b - 1) / b;
}

// streaming (read-once / write-once) global accesses: keep the 126 MB L2 for data that is re-read (halos, twiddles)
__device__ __forceinline__ float4 ldStream4(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldStream2(const float2* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void stStream4(float4* p, float4 v) { asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }
__device__ __forceinline__ void stStream2(float2* p, float2 v) { asm volatile("st.global.L1::no_allocate.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory"); }

// cos/sin of the mixer phase (Rotator.hpp:59-60 calls std::cos / std::sin). The reference's phase lives in [0, 2 pi]
// after the first wrap, so the common case is a branch-free three-term Cody-Waite reduction by pi/2 (exact products via
// FMA, quotient from the round-to-nearest magic constant) and the Cephes single-precision minimax polynomials on
// [-pi/4, pi/4] (sin: degree 7, cos: degree 8); checked against the oracle within the mixer tolerance
// (tests/test_gpu_parity.py). Anything outside |x| <= 64 (a user-set start phase far from [0, 2 pi]) takes the library.
static __device__ __noinline__ void sinCosLibrary(float x, float* s, float* c) { sincosf(x, s, c); }
constexpr float kMixerFastRange = 64.f; // mixerSinCosFast is used for |x| <= this
__device__ __forceinline__ void mixerSinCosFast(float x, float* s, float* c) {
    const float    magic = 12582912.f;                                  // 1.5 * 2^23: adding it rounds to an integer
    const float    t     = fmaf(x, 0.636619747f, magic);                // x * 2/pi, rounded to nearest integer
    const unsigned q     = __float_as_uint(t);                          // low bits = quadrant (two's complement for t < magic)
    const float    qf    = __fsub_rn(t, magic);
    float          r     = fmaf(qf, -1.57079601e+00f, x);               // pi/2 split in three: 24 + 24 + 24 bits
    r                    = fmaf(qf, -3.13916473e-07f, r);
    r                    = fmaf(qf, -5.39030253e-15f, r);
    const float r2       = r * r;
    float       sp       = fmaf(-1.95152959e-4f, r2, 8.33216087e-3f);
    sp                   = fmaf(sp, r2, -1.66666546e-1f);
    const float sinR     = fmaf(sp * r2, r, r);
    float       cp       = fmaf(2.44331571e-5f, r2, -1.38873163e-3f);
    cp                   = fmaf(cp, r2, 4.16666457e-2f);
    cp                   = fmaf(cp, r2, -0.5f);
    const float cosR     = fmaf(cp, r2, 1.f);
    const bool  swap     = (q & 1u) != 0;
    float       sv       = swap ? cosR : sinR;
    float       cv       = swap ? sinR : cosR;
    sv                   = (q & 2u) != 0 ? -sv : sv;
    cv                   = ((q + 1u) & 2u) != 0 ? -cv : cv;
    *s                   = sv;
    *c                   = cv;
}
// The same operation sequence on TWO phases at once, lane by lane in packed f32x2 arithmetic (fma / mul / add .rn per
// half are the scalar IEEE operations): bit-identical to two calls of mixerSinCosFast, about half the issue slots --
// the mixer kernels are bound by instruction issue, not by HBM or the fp32 pipe.
__device__ __forceinline__ void mixerSinCosFast2(float x0, float x1, float* s0, float* c0, float* s1, float* c1) {
    using P = unsigned long long;
    auto pk  = [](float a, float b) { P r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; };
    auto sp1 = [&](float a) { return pk(a, a); };
    auto fma = [](P a, P b, P c) { P r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; };
    auto mul = [](P a, P b) { P r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; };
    auto add = [](P a, P b) { P r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; };
    auto lo  = [](P v) { return __uint_as_float(static_cast<unsigned>(v)); };
    auto hi  = [](P v) { return __uint_as_float(static_cast<unsigned>(v >> 32)); };
    const float magic = 12582912.f;
    const P     x     = pk(x0, x1);
    const P     t     = fma(x, sp1(0.636619747f), sp1(magic));
    const P     qf    = add(t, sp1(-magic)); // t - magic, exact like the scalar __fsub_rn
    P           r     = fma(qf, sp1(-1.57079601e+00f), x);
    r                 = fma(qf, sp1(-3.13916473e-07f), r);
    r                 = fma(qf, sp1(-5.39030253e-15f), r);
    const P r2        = mul(r, r);
    P       sp        = fma(sp1(-1.95152959e-4f), r2, sp1(8.33216087e-3f));
    sp                = fma(sp, r2, sp1(-1.66666546e-1f));
    const P sinR      = fma(mul(sp, r2), r, r);
    P       cp        = fma(sp1(2.44331571e-5f), r2, sp1(-1.38873163e-3f));
    cp                = fma(cp, r2, sp1(4.16666457e-2f));
    cp                = fma(cp, r2, sp1(-0.5f));
    const P cosR      = fma(cp, r2, sp1(1.f));
    const unsigned q0 = __float_as_uint(lo(t)), q1 = __float_as_uint(hi(t));
    float sv0 = (q0 & 1u) != 0 ? lo(cosR) : lo(sinR), cv0 = (q0 & 1u) != 0 ? lo(sinR) : lo(cosR);
    float sv1 = (q1 & 1u) != 0 ? hi(cosR) : hi(sinR), cv1 = (q1 & 1u) != 0 ? hi(sinR) : hi(cosR);
    *s0 = (q0 & 2u) != 0 ? -sv0 : sv0;
    *c0 = ((q0 + 1u) & 2u) != 0 ? -cv0 : cv0;
    *s1 = (q1 & 2u) != 0 ? -sv1 : sv1;
    *c1 = ((q1 + 1u) & 2u) != 0 ? -cv1 : cv1;
}
__device__ __forceinline__ void mixerSinCos(float x, float* s, float* c) {
    if (!(fabsf(x) <= kMixerFastRange)) {
        sinCosLibrary(x, s, c);
        return;
    }
    mixerSinCosFast(x, s, c);
}

// std::complex<float> product with the reference's rounding: libgcc __mulsc3 = separately rounded products, then
// subtract / add, then C99 Annex G recovery when both parts come out NaN.
static __device__ __noinline__ float2 complexMulRecover(float a, float b, float c, float d, float x, float y) {
    const float ac = __fmul_rn(a, c), bd = __fmul_rn(b, d), ad = __fmul_rn(a, d), bc = __fmul_rn(b, c);
    {
        bool recalc = false;
        if (isinf(a) || isinf(b)) {
            a = copysignf(isinf(a) ? 1.f : 0.f, a);
            b = copysignf(isinf(b) ? 1.f : 0.f, b);
            if (isnan(c)) c = copysignf(0.f, c);
            if (isnan(d)) d = copysignf(0.f, d);
            recalc = true;
        }
        if (isinf(c) || isinf(d)) {
            c = copysignf(isinf(c) ? 1.f : 0.f, c);
            d = copysignf(isinf(d) ? 1.f : 0.f, d);
            if (isnan(a)) a = copysignf(0.f, a);
            if (isnan(b)) b = copysignf(0.f, b);
            recalc = true;
        }
        if (!recalc && (isinf(ac) || isinf(bd) || isinf(ad) || isinf(bc))) {
            if (isnan(a)) a = copysignf(0.f, a);
            if (isnan(b)) b = copysignf(0.f, b);
            if (isnan(c)) c = copysignf(0.f, c);
            if (isnan(d)) d = copysignf(0.f, d);
            recalc = true;
        }
        if (recalc) {
            x = __fmul_rn(INFINITY, __fsub_rn(__fmul_rn(a, c), __fmul_rn(b, d)));
            y = __fmul_rn(INFINITY, __fadd_rn(__fmul_rn(a, d), __fmul_rn(b, c)));
        }
    }
    return make_float2(x, y);
}
// Packed form: (ac, bc) and (ad, bd) are two FMUL2 with a broadcast operand, (ac - bd, bc + ad) is ONE FFMA2 of the
// swapped second pair against (-1, +1): an fma with a multiplier of +-1 rounds once, exactly like the subtraction /
// addition of the separately rounded products, and ptxas keeps the products apart (it has no second multiply to fuse).
// Three instructions per sample instead of six; bit-identical (tests/test_gpu_parity.py, incl. inf / nan / signed zeros).
__device__ __forceinline__ float2 complexMulAnnexG(float a, float b, float c, float d) {
    unsigned long long ab, cc, dd, p1, p2, p2s, pm, r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ab) : "f"(a), "f"(b));
    asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c));
    asm("mov.b64 %0, {%1, %1};" : "=l"(dd) : "f"(d));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p1) : "l"(ab), "l"(cc)); // (ac, bc)
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p2) : "l"(ab), "l"(dd)); // (ad, bd)
    float ad, bd;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(ad), "=f"(bd) : "l"(p2));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p2s) : "f"(bd), "f"(ad));
    asm("mov.b64 %0, {%1, %2};" : "=l"(pm) : "f"(-1.f), "f"(1.f));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(p2s), "l"(pm), "l"(p1)); // (ac - bd, bc + ad)
    float x, y;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(r));
    if (x != x && y != y) { // both parts NaN: rare, kept out of line
        return complexMulRecover(a, b, c, d, x, y);
    }
    return make_float2(x, y);
}

// std::complex<float> quotient with the reference's rounding: libgcc __divsc3 evaluates in double when the host has
// hardware double (x86-64), rounds once to float, then applies Annex G recovery.
__device__ __forceinline__ float2 complexDivAnnexG(float a, float b, float c, float d) {
    const double aa = a, bb = b, cc = c, dd = d;
    const double denom = __dadd_rn(__dmul_rn(cc, cc), __dmul_rn(dd, dd));
    float        x     = __double2float_rn(__ddiv_rn(__dadd_rn(__dmul_rn(aa, cc), __dmul_rn(bb, dd)), denom));
    float        y     = __double2float_rn(__ddiv_rn(__dsub_rn(__dmul_rn(bb, cc), __dmul_rn(aa, dd)), denom));
    if (isnan(x) && isnan(y)) {
        if (c == 0.f && d == 0.f && (!isnan(a) || !isnan(b))) {
            x = __fmul_rn(copysignf(INFINITY, c), a);
            y = __fmul_rn(copysignf(INFINITY, c), b);
        } else if ((isinf(a) || isinf(b)) && isfinite(c) && isfinite(d)) {
            a = copysignf(isinf(a) ? 1.f : 0.f, a);
            b = copysignf(isinf(b) ? 1.f : 0.f, b);
            x = __fmul_rn(INFINITY, __fadd_rn(__fmul_rn(a, c), __fmul_rn(b, d)));
            y = __fmul_rn(INFINITY, __fsub_rn(__fmul_rn(b, c), __fmul_rn(a, d)));
        } else if ((isinf(c) || isinf(d)) && isfinite(a) && isfinite(b)) {
            c = copysignf(isinf(c) ? 1.f : 0.f, c);
            d = copysignf(isinf(d) ? 1.f : 0.f, d);
            x = __fmul_rn(0.f, __fadd_rn(__fmul_rn(a, c), __fmul_rn(b, d)));
            y = __fmul_rn(0.f, __fsub_rn(__fmul_rn(b, c), __fmul_rn(a, d)));
        }
    }
    return make_float2(x, y);
}

} // namespace gr4b200
