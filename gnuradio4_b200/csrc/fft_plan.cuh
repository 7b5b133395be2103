// The plan behind gr4b200_fft_plan_create (fft.cu), also read by the fused FIR -> FFT step (fir_fft.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>

struct gr4b200_fft_plan {
    int                  device  = 0;       // the device the plan's memory lives on
    size_t               n       = 0;
    float*               windowT = nullptr; // device: window in the per-thread layout of pass 1, or nullptr
    float2*              tables  = nullptr; // device: twiddle tables of all passes
    bool                 useTma  = true;    // GR4B200_FFT_TMA=0 forces the direct-load variant (A/B timing)
    // n > 8192 (fft_large.cuh): n = n1 * n2, two passes of column transforms through a scratch buffer
    size_t               n1 = 0, n2 = 0;
    float2*              tables1     = nullptr; // pass tables of the n1- and n2-point column transforms
    float2*              tables2     = nullptr;
    float2*              twiddleN    = nullptr; // W_n^j, j in [0, n)
    float*               windowN     = nullptr; // window in natural order, or nullptr
    float2*              scratch     = nullptr; // intermediate A[n2][n1] of one slice of transforms (grows on demand)
    size_t               scratchSize = 0;       // in complex samples
    // n not a power of two, or below 16 (fft_bluestein.cuh): chirp-z through an inner plan of bluesteinM points
    size_t               bluesteinM    = 0;       // 0: the radix kernels serve n directly
    gr4b200_fft_plan*    inner         = nullptr; // the bluesteinM-point plan
    float2*              chirpConj     = nullptr; // e^{-j pi i^2 / n}, i in [0, n)
    float2*              chirpSpectrum = nullptr; // FFT_M of the symmetric chirp, divided by M
    float2*              work          = nullptr; // two arrays of M points per transform of one slice (grows on demand)
    size_t               workSize      = 0;
    float2*              spectrum      = nullptr; // block mode: the spectrum before the plane kernel (grows on demand)
    size_t               spectrumSize  = 0;
};

