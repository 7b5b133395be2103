// Shared helpers for the gr4b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <string>

#include "../../include/gr4b200.h"
#include "sincos_core.cuh"

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

// launch-error check: every kernel launch in the C ABI ends with this (cudaGetLastError -> work::Status::ERROR); it also
// counts the launch (gr4b200_launch_count)
std::atomic<unsigned long long>& launchCounter(); // runtime.cu
inline int checkLaunch(const char* kernelName, unsigned launches = 1) {
    launchCounter().fetch_add(launches, std::memory_order_relaxed);
    return checkCuda(cudaGetLastError(), kernelName);
}

inline cudaStream_t asStream(void* stream) { return static_cast<cudaStream_t>(stream); }

// ---- programmatic dependent launch ---------------------------------------------------------------------------------------
// A streaming flowgraph issues FIR, FFT block, FIR, FFT block, ... on one stream, each kernel depending on its predecessor;
// at work chunks of 2^16 .. 2^20 samples the gaps between dependent launches are a third of the time. The
// kernels of that chain therefore (a) let their successor be launched as soon as all of their own CTAs have started
// (gridDependencyLaunch, first instruction) and (b) wait for their predecessor to have finished and flushed
// (gridDependencyWait) only in front of their first access to stream data -- shared-memory set-up, barrier
// initialisation and constant tables overlap the predecessor's tail. Only kernels that contain the wait are launched with
// the attribute; everything else keeps the plain stream order. GR4B200_PDL=0 switches the attribute off (A/B timing).
#ifdef __CUDACC__
__device__ __forceinline__ void gridDependencyLaunch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void gridDependencyWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
inline bool pdlEnabled() {
    static const bool enabled = [] { const char* e = std::getenv("GR4B200_PDL"); return e == nullptr || e[0] != '0'; }();
    return enabled;
}
template<typename... Params, typename... Args>
inline cudaError_t launchDependent(void (*kernel)(Params...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t config = {};
    config.gridDim          = grid;
    config.blockDim         = block;
    config.dynamicSmemBytes = smem;
    config.stream           = stream;
    cudaLaunchAttribute attribute[1];
    attribute[0].id                                         = cudaLaunchAttributeProgrammaticStreamSerialization;
    attribute[0].val.programmaticStreamSerializationAllowed = pdlEnabled() ? 1 : 0;
    config.attrs                                            = attribute;
    config.numAttrs                                         = 1;
    return cudaLaunchKernelEx(&config, kernel, std::forward<Args>(args)...);
}
#endif

int smCount(); // SMs of the current device (cached per device)

// Plans own device memory: they record the device they were created on, and every compute entry checks that the caller
// is on that device (a launch from another current device would dereference foreign pointers).
inline int currentDevice() {
    int device = -1;
    if (cudaGetDevice(&device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return device;
}
inline int checkPlanDevice(int planDevice, const char* what) {
    const int device = currentDevice();
    if (device == planDevice) {
        return GR4B200_OK;
    }
    return fail(std::string(what) + ": plan lives on cuda:" + std::to_string(planDevice) + " but the calling thread's current device is cuda:" + std::to_string(device) + " (cudaSetDevice / gr4b200_init first)");
}

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

// cos/sin of the mixer phase (Rotator.hpp:59-60 calls std::cos / std::sin): the C library's own operation sequence,
// restated in sincos_core.cuh and evaluated on the FP64 pipe => the library's bits for every float argument. The
// reference's phase lives in [0, 2 pi] after the first wrap; |x| < 120 is straight-line code, anything else (a user-set
// start phase far away, inf, NaN) is kept out of line.
static __device__ __noinline__ void sinCosOutOfLine(float x, float* s, float* c) { sinCosGlibc(x, s, c); }
constexpr float kMixerFastRange = 64.f; // callers that advance a phase by up to 16 steps of at most 3.5 rad test against this
__device__ __forceinline__ void mixerSinCosFast(float x, float* s, float* c) { sinCosGlibcSmall(x, s, c); } // |x| < 120
__device__ __forceinline__ void mixerSinCosInRange(float x, float* s, float* c) { sinCosGlibcSmall<false>(x, s, c); } // 0 <= x < 120, not -0
__device__ __forceinline__ void mixerSinCos(float x, float* s, float* c) {
    if (!(fabsf(x) < kSinCosSmallLimit)) {
        sinCosOutOfLine(x, s, c);
        return;
    }
    sinCosGlibcSmall(x, s, c);
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
