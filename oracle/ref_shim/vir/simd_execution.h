// TEST INFRASTRUCTURE ONLY -- see vir/simd.h. The reference FFT includes this header but the C2C/R2C
// transform path never calls into it (SimdFFT.hpp:86), so an empty stand-in suffices.
#ifndef GR4B200_ORACLE_SHIM_VIR_SIMD_EXECUTION_H
#define GR4B200_ORACLE_SHIM_VIR_SIMD_EXECUTION_H
#endif
