// Micro-benchmark (GPU box only): issue rate of packed fp32x2 math (FFMA2/FMUL2/FADD2) vs scalar FFMA/FMUL+FADD.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_f32x2 scripts/ubench_f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kChains = 16;

__global__ void scalarFma(float* out, float a, float b) {
    float acc[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc[c] = threadIdx.x + c;
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) acc[c] = __fmaf_rn(acc[c], a, b);
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void scalarMulAdd(float* out, float a, float b) {
    float acc[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc[c] = threadIdx.x + c;
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) acc[c] = __fadd_rn(__fmul_rn(acc[c], a), b);
    }
    float s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long pack(float lo, float hi) { return (static_cast<unsigned long long>(__float_as_uint(hi)) << 32) | __float_as_uint(lo); }

__global__ void packedFma(float* out, float a, float b) {
    unsigned long long acc[kChains];
    const unsigned long long va = pack(a, a), vb = pack(b, b);
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc[c] = pack(threadIdx.x + c, threadIdx.x - c);
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc[c]) : "l"(acc[c]), "l"(va), "l"(vb));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s ^= acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(static_cast<unsigned>(s) ^ static_cast<unsigned>(s >> 32));
}

__global__ void packedMulAdd(float* out, float a, float b) {
    unsigned long long acc[kChains];
    const unsigned long long va = pack(a, a), vb = pack(b, b);
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc[c] = pack(threadIdx.x + c, threadIdx.x - c);
    for (int i = 0; i < kIters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
            unsigned long long t;
            asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(acc[c]), "l"(va));
            asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(acc[c]) : "l"(t), "l"(vb));
        }
    }
    unsigned long long s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s ^= acc[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(static_cast<unsigned>(s) ^ static_cast<unsigned>(s >> 32));
}

template<typename K>
float timeKernel(K kernel, float* out, int grid, int block) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    kernel<<<grid, block>>>(out, 0.999f, 0.001f);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) kernel<<<grid, block>>>(out, 0.999f, 0.001f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 8, block = 256;
    float*    out;
    cudaMalloc(&out, sizeof(float) * grid * block);
    const double lanesOps = double(grid) * block * kIters * kChains; // per-thread math instructions of the scalar FMA kernel
    float ms;
    ms = timeKernel(scalarFma, out, grid, block);
    printf("scalar FFMA      : %.3f ms  %.2f T lane-instr/s  (%.1f TFLOP/s)\n", ms, lanesOps / ms / 1e9, 2 * lanesOps / ms / 1e9);
    ms = timeKernel(scalarMulAdd, out, grid, block);
    printf("scalar FMUL+FADD : %.3f ms  %.2f T lane-instr/s  (mul-add pairs/s %.2f T)\n", ms, 2 * lanesOps / ms / 1e9, lanesOps / ms / 1e9);
    ms = timeKernel(packedFma, out, grid, block);
    printf("packed FFMA2     : %.3f ms  %.2f T lane-instr/s  (%.1f TFLOP/s)\n", ms, lanesOps / ms / 1e9, 4 * lanesOps / ms / 1e9);
    ms = timeKernel(packedMulAdd, out, grid, block);
    printf("packed FMUL2+FADD2: %.3f ms  %.2f T lane-instr/s  (mul-add pairs/s %.2f T)\n", ms, 2 * lanesOps / ms / 1e9, 2 * lanesOps / ms / 1e9);
    return 0;
}
