// cos/sin of the mixer phase with the reference's bits.
//
// Rotator.hpp:59-60 calls std::cos / std::sin on a float, i.e. glibc's cosf / sinf (GCC may merge the pair into one
// sincosf call; all three share one kernel, sysdeps/ieee754/flt-32/sincosf.h, and return identical bits). That kernel
// is not a float computation: the argument is widened to double, reduced by pi/2 in double (one multiplication by
// 2/pi * 2^24, an integer shift for the quadrant, ONE fused multiply-subtract for the remainder), two short double
// polynomials are evaluated and the result is rounded to float once. This header restates that arithmetic -- the
// x86-64 build glibc selects on every machine with FMA + AVX2 (__sincosf_fma: each `a + b * c` of the source is one
// vfmadd, checked against the disassembly of libm 2.39) -- so the device result is the library's result, bit for bit,
// for every float: scripts/verify_sincos_all_floats.cu sweeps all 2^32 arguments against libm on the host (both the
// library's own sequence and the shorter one the device runs), tests/test_host_emulation.py runs the [0, 2 pi] band and a
// sample of the rest in the CPU suite.
// On the device the arithmetic runs on the FP64 pipe (DMUL / DFMA), which the mixer and the fused DDC leave idle.
#pragma once

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>

#ifndef GR4B200_HD
#define GR4B200_HD __host__ __device__ __forceinline__
#endif

namespace gr4b200 {
namespace sincos_detail {

GR4B200_HD unsigned floatBits(float v) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(v);
#else
    unsigned u;
    std::memcpy(&u, &v, sizeof u);
    return u;
#endif
}
GR4B200_HD float negateIf(float v, bool negate) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(__float_as_uint(v) ^ (negate ? 0x80000000u : 0u));
#else
    return negate ? -v : v;
#endif
}
GR4B200_HD double fmaD(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
GR4B200_HD double mulD(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}

// polynomial coefficients of the library's table (quadrants 0/1; quadrants 2/3 use the negated cosine set, which
// negates the cosine result exactly, so one set and a sign flip give the same bits)
constexpr double kC0 = 0x1p0, kC1 = -0x1.ffffffd0c621cp-2, kC2 = 0x1.55553e1068f19p-5, kC3 = -0x1.6c087e89a359dp-10, kC4 = 0x1.99343027bf8c3p-16;
constexpr double kS1 = -0x1.555545995a603p-3, kS2 = 0x1.1107605230bc4p-7, kS3 = -0x1.994eb3774cf24p-13;
constexpr double kHalfPiInv24 = 0x1.45F306DC9C883p+23; // 2/pi * 2^24
constexpr double kHalfPi      = 0x1.921FB54442D18p0;
constexpr double kPi63        = 0x1.921FB54442D18p-62; // pi/2 * 2^-62: scale of the 62-bit fixed-point remainder

// sine polynomial on the reduced argument x (x2 = x*x), cosine polynomial on x2; both rounded to float once
GR4B200_HD void polynomials(double x, double x2, float* sinPoly, float* cosPoly) {
    const double x3 = mulD(x2, x);
    const double x4 = mulD(x2, x2);
    const double c2 = fmaD(x2, kC4, kC3);
    const double s1 = fmaD(x2, kS3, kS2);
    const double c1 = fmaD(x2, kC1, kC0);
    const double x5 = mulD(x3, x2);
    const double x6 = mulD(x4, x2);
    const double s  = fmaD(x3, kS1, x);
    const double c  = fmaD(x4, kC2, c1);
#ifdef __CUDA_ARCH__
    *sinPoly = __double2float_rn(fmaD(x5, s1, s));
    *cosPoly = __double2float_rn(fmaD(x6, c2, c));
#else
    *sinPoly = static_cast<float>(fmaD(x5, s1, s));
    *cosPoly = static_cast<float>(fmaD(x6, c2, c));
#endif
}

// |y| >= 120: the library multiplies the 24-bit mantissa by 96 bits of 4/pi picked by the exponent (integer arithmetic)
GR4B200_HD double reduceLarge(unsigned xi, int* quadrant) {
    constexpr unsigned kInvPio4[24] = {0xa2, 0xa2f9, 0xa2f983, 0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529, 0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0, 0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};
    const unsigned     first        = (xi >> 26) & 15u;
    const int          shift        = static_cast<int>((xi >> 23) & 7u);
    unsigned           m            = ((xi & 0xffffffu) | 0x800000u) << shift;
    unsigned long long res0         = static_cast<unsigned long long>(static_cast<unsigned>(m * kInvPio4[first])); // 32-bit product, as the library computes it
    const unsigned long long res1   = static_cast<unsigned long long>(m) * kInvPio4[first + 4];
    const unsigned long long res2   = static_cast<unsigned long long>(m) * kInvPio4[first + 8];
    res0                            = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    const unsigned long long n = (res0 + (1ull << 61)) >> 62;
    res0 -= n << 62;
    *quadrant = static_cast<int>(n);
#ifdef __CUDA_ARCH__
    return mulD(__ll2double_rn(static_cast<long long>(res0)), kPi63);
#else
    return static_cast<double>(static_cast<long long>(res0)) * kPi63;
#endif
}

} // namespace sincos_detail

// sinf(y) and cosf(y) as glibc 2.39 (x86-64, FMA build) returns them, for |y| < 120 (the caller checks): the form the
// device runs. It is NOT the library's operation sequence (sinCosGlibcReference below is) but a shorter one that
// returns the same bits for every one of the 2^31.06 float arguments in range -- established by exhaustive comparison on
// the host (scripts/verify_sincos_all_floats.cu, tests/test_host_emulation.py), not by argument:
//  * quadrant: one fused product x * 2/pi + 1.5 * 2^52 leaves n in the low mantissa bits and n as a double after
//    subtracting the constant again (the library multiplies by 2/pi * 2^24, truncates to int, adds 2^23, shifts, and
//    converts back: three conversions on a pipe that issues them at a fraction of the DFMA rate);
//  * the two polynomials in Horner form: 9 instead of 12 double operations; the last-bit differences in double never
//    cross a float rounding boundary for any argument;
//  * the library's special cases |y| < 2^-12 -> (y, 1) and |y| < pi/4 -> no reduction are the n = 0 case of the
//    reduction (fma(-0, pi/2, x) = x exactly) except for the sign of sin(-0), which a select restores: threads of a warp
//    never diverge here;
//  * the quadrant's sign flips and the sine / cosine swap are bit operations on the rounded results.
// KeepNegativeZero = false: the caller knows y is not -0 (a mixer phase in [0, 2 pi_f] after a non-zero step never is:
// an exact zero sum rounds to +0) and the select that restores sin(-0) = -0 drops out.
template<bool KeepNegativeZero = true>
GR4B200_HD void sinCosGlibcSmall(float y, float* sinOut, float* cosOut) {
    using namespace sincos_detail;
    constexpr double kRoundMagic = 0x1.8p52;                // adding it rounds a double in (-2^51, 2^51) to an integer
    constexpr double kTwoOverPi  = 0x1.45F306DC9C883p-1;    // kHalfPiInv24 * 2^-24
    const double     x           = static_cast<double>(y);  // one F2F on the device, at the fp32 rate
    const double     t           = fmaD(x, kTwoOverPi, kRoundMagic);
#ifdef __CUDA_ARCH__
    const unsigned n  = static_cast<unsigned>(__double2loint(t));
    const double   nd = __dadd_rn(t, -kRoundMagic);
#else
    long long tBits;
    std::memcpy(&tBits, &t, sizeof tBits);
    const unsigned n  = static_cast<unsigned>(tBits);
    const double   nd = t - kRoundMagic;
#endif
    const double xr = fmaD(-nd, kHalfPi, x);
    const double x2 = mulD(xr, xr);
    const double x3 = mulD(x2, xr);
    double       ps = fmaD(x2, kS3, kS2);
    ps              = fmaD(x2, ps, kS1);
    double pc       = fmaD(x2, kC4, kC3);
    pc              = fmaD(x2, pc, kC2);
    pc              = fmaD(x2, pc, kC1);
#ifdef __CUDA_ARCH__
    const unsigned sp = __float_as_uint(__double2float_rn(fmaD(x3, ps, xr)));
    const unsigned cp = __float_as_uint(__double2float_rn(fmaD(x2, pc, kC0)));
#else
    const unsigned sp = floatBits(static_cast<float>(fmaD(x3, ps, xr)));
    const unsigned cp = floatBits(static_cast<float>(fmaD(x2, pc, kC0)));
#endif
    // quadrant signs as bit operations on the float patterns; odd quadrants swap sine and cosine
    // (the same truth table as "negate, then swap" with two instructions fewer: select first, then sin(x + n pi/2) is negative
    // for n in {2, 3} and cos for n in {1, 2}: bit 1 of n and of n + 1, moved to the sign position by one shared shift)
    const bool     swap    = (n & 1u) != 0;
    const unsigned n30     = n << 30; // bit 31 = bit 1 of n, bit 30 = bit 0 of n
    const unsigned sOut    = (swap ? cp : sp) ^ (n30 & 0x80000000u);
    const unsigned cOut    = (swap ? sp : cp) ^ ((n30 + 0x40000000u) & 0x80000000u);
    // |y| < 2^-12: the library returns (y, 1). The formulas above already give exactly that (n = 0; the corrections are
    // below half an ulp of y and of 1) except for sin(-0), whose sign the fused sum loses: a select on the sine alone
#ifdef __CUDA_ARCH__
    *sinOut = KeepNegativeZero && fabsf(y) < 0x1p-12f ? y : __uint_as_float(sOut);
    *cosOut = __uint_as_float(cOut);
#else
    float sv, cv;
    std::memcpy(&sv, &sOut, sizeof sv);
    std::memcpy(&cv, &cOut, sizeof cv);
    *sinOut = KeepNegativeZero && std::fabs(y) < 0x1p-12f ? y : sv;
    *cosOut = cv;
#endif
}

// The library's own operation sequence for |y| < 120 (reduce_fast + sincosf_poly of sysdeps/ieee754/flt-32/sincosf.h as
// the FMA build compiles them): kept as the statement of what sinCosGlibcSmall has to equal.
GR4B200_HD void sinCosGlibcReference(float y, float* sinOut, float* cosOut) {
    using namespace sincos_detail;
    const unsigned top = (floatBits(y) >> 20) & 0x7ffu;
    const double   x   = static_cast<double>(y);
    if (top < 0x3f4u) { // |y| < pi/4
        if (top < 0x398u) {
            *sinOut = y;
            *cosOut = 1.f;
            return;
        }
        polynomials(x, mulD(x, x), sinOut, cosOut);
        return;
    }
    const double r = mulD(x, kHalfPiInv24);
#ifdef __CUDA_ARCH__
    const int    n  = (__double2int_rz(r) + 0x800000) >> 24;
    const double xr = fmaD(-__int2double_rn(n), kHalfPi, x);
#else
    const int    n  = (static_cast<int>(r) + 0x800000) >> 24;
    const double xr = fmaD(-static_cast<double>(n), kHalfPi, x);
#endif
    float sp, cp;
    polynomials(xr, mulD(xr, xr), &sp, &cp);
    sp              = negateIf(sp, (((n >> 1) ^ n) & 1) != 0);
    cp              = negateIf(cp, (n & 2) != 0);
    const bool swap = (n & 1) != 0;
    *sinOut         = swap ? cp : sp;
    *cosOut         = swap ? sp : cp;
}

constexpr float kSinCosSmallLimit = 120.f; // sinCosGlibcSmall covers |y| < this (bit pattern test: top 12 bits < 0x42f)

// any float
GR4B200_HD void sinCosGlibc(float y, float* sinOut, float* cosOut) {
    using namespace sincos_detail;
    const unsigned xi  = floatBits(y);
    const unsigned top = (xi >> 20) & 0x7ffu;
    if (top < 0x42fu) {
        sinCosGlibcSmall(y, sinOut, cosOut);
        return;
    }
    if (top >= 0x7f8u) { // inf or NaN
        *sinOut = *cosOut = y - y;
        return;
    }
    int          n;
    const double xr        = reduceLarge(xi, &n);
    const int    signIndex = n + static_cast<int>(xi >> 31);
    float        sp, cp;
    polynomials(xr, mulD(xr, xr), &sp, &cp);
    sp              = negateIf(sp, (((signIndex >> 1) ^ signIndex) & 1) != 0);
    cp              = negateIf(cp, (signIndex & 2) != 0);
    const bool swap = (n & 1) != 0;
    *sinOut         = swap ? cp : sp;
    *cosOut         = swap ? sp : cp;
}

} // namespace gr4b200
