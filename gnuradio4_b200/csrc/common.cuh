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

// std::complex<float> product with the reference's rounding: libgcc __mulsc3 = separately rounded products, then
// subtract / add, then C99 Annex G recovery when both parts come out NaN.
__device__ __forceinline__ float2 complexMulAnnexG(float a, float b, float c, float d) {
    const float ac = __fmul_rn(a, c), bd = __fmul_rn(b, d), ad = __fmul_rn(a, d), bc = __fmul_rn(b, c);
    float       x = __fsub_rn(ac, bd), y = __fadd_rn(ad, bc);
    if (isnan(x) && isnan(y)) {
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
