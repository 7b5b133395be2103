// Micro-benchmark (GPU box only): throughput of the FP64 pipe and of the conversions the mixer's library-exact sin/cos
// needs (csrc/sincos_core.cuh), in thread-instructions per clock per SM, and the rate of the whole function.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/ubench_fp64 scripts/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "../gnuradio4_b200/csrc/sincos_core.cuh"

constexpr int kIters  = 2048;
constexpr int kChains = 8;

template<int Op>
__global__ void opKernel(float* out, float fa, double da, double db) {
    double d[kChains];
    float  f[kChains];
    int    n[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        d[c] = threadIdx.x * 1e-3 + c;
        f[c] = threadIdx.x * 1e-3f + c;
        n[c] = threadIdx.x + c;
    }
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
            if constexpr (Op == 0) { // DFMA
                d[c] = __fma_rn(d[c], da, db);
            } else if constexpr (Op == 1) { // DMUL
                d[c] = __dmul_rn(d[c], da);
            } else if constexpr (Op == 2) { // DADD
                d[c] = __dadd_rn(d[c], db);
            } else if constexpr (Op == 3) { // f32 -> f64 (+ an fp32 add to keep the chain in float)
                asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[c]) : "f"(f[c]));
                f[c] = __fadd_rn(f[c], fa);
            } else if constexpr (Op == 4) { // f64 -> f32
                asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[c]) : "d"(d[c]));
                d[c] = __longlong_as_double(__double_as_longlong(d[c]) ^ __float_as_int(f[c]));
            } else if constexpr (Op == 5) { // f64 -> s32 (truncate)
                asm volatile("cvt.rzi.s32.f64 %0, %1;" : "=r"(n[c]) : "d"(d[c]));
                d[c] = __longlong_as_double(__double_as_longlong(d[c]) ^ (n[c] & 1));
            } else if constexpr (Op == 6) { // s32 -> f64
                asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d[c]) : "r"(n[c]));
                n[c] += __double2hiint(d[c]) & 1;
            } else if constexpr (Op == 7) { // FFMA (reference rate)
                f[c] = __fmaf_rn(f[c], fa, 0.001f);
            } else if constexpr (Op == 8) { // DFMA + 2 FFMA interleaved: do the pipes overlap?
                d[c] = __fma_rn(d[c], da, db);
                f[c] = __fmaf_rn(f[c], fa, 0.001f);
                f[c] = __fmaf_rn(f[c], fa, 0.002f);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += d[c] + f[c] + n[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = static_cast<float>(s);
}

__global__ void sinCosKernel(float* out, float x0, float step) {
    float x = x0 + threadIdx.x * step, acc = 0.f;
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
            float s, co;
            gr4b200::sinCosGlibcSmall(x, &s, &co);
            acc += s * co;
            x += step;
            x = x > 6.2831853f ? x - 6.2831853f : x;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template<typename L>
float timeIt(L launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 3; ++r) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms / 3;
}

int main() {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int grid = sms * 8, block = 256;
    float*    out;
    cudaMalloc(&out, sizeof(float) * grid * block);
    const double ops   = double(grid) * block * kIters * kChains;
    const double clkHz = khz * 1e3;
    const char*  names[] = {"DFMA", "DMUL", "DADD", "F2F.F64.F32 (+FADD)", "F2F.F32.F64 (+LOP)", "F2I.S32.F64 (+LOP)", "I2F.F64.S32 (+IADD)", "FFMA", "DFMA + 2 FFMA"};
    auto report = [&](int op, float ms) { printf("{\"op\": \"%s\", \"ms\": %.3f, \"T_per_s\": %.2f, \"per_clk_per_sm\": %.1f}\n", names[op], ms, ops / ms / 1e9, ops / (ms * 1e-3) / clkHz / sms); };
    report(0, timeIt([&] { opKernel<0><<<grid, block>>>(out, 0.999f, 0.999, 0.001); }));
    report(1, timeIt([&] { opKernel<1><<<grid, block>>>(out, 0.999f, 0.999, 0.001); }));
    report(2, timeIt([&] { opKernel<2><<<grid, block>>>(out, 0.999f, 0.999, 0.001); }));
    report(3, timeIt([&] { opKernel<3><<<grid, block>>>(out, 0.999f, 0.999, 0.001); }));
    report(4, timeIt([&] { opKernel<4><<<grid, block>>>(out, 0.999f, 0.999, 0.001); }));
    report(5, timeIt([&] { opKernel<5><<<grid, block>>>(out, 0.999f, 0.999, 0.001); }));
    report(6, timeIt([&] { opKernel<6><<<grid, block>>>(out, 0.999f, 0.999, 0.001); }));
    report(7, timeIt([&] { opKernel<7><<<grid, block>>>(out, 0.999f, 0.999, 0.001); }));
    report(8, timeIt([&] { opKernel<8><<<grid, block>>>(out, 0.999f, 0.999, 0.001); }));
    const float ms = timeIt([&] { sinCosKernel<<<grid, block>>>(out, 0.1f, 0.01f); });
    printf("{\"op\": \"sinCosGlibcSmall\", \"ms\": %.3f, \"G_calls_per_s\": %.1f, \"clk_per_call_per_sm\": %.3f, \"sm_clock_mhz\": %.0f}\n", ms, ops / ms / 1e6, (ms * 1e-3) * clkHz * sms / ops, clkHz / 1e6);
    return 0;
}
