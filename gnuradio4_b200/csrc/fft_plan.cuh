// The plan behind gr4b200_fft_plan_create (fft.cu), also read by the fused FIR -> FFT step (fir_fft.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>

struct gr4b200_fft_plan {
    size_t               n       = 0;
    float*               windowT = nullptr; // device: window in the per-thread layout of pass 1, or nullptr
    float2*              tables  = nullptr; // device: twiddle tables of all passes
    bool                 useTma  = true;    // GR4B200_FFT_TMA=0 forces the direct-load variant (A/B timing)
};

