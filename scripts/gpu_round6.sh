#!/bin/bash
# exact-FIR arithmetic forms A/B (GR4B200_FIR_EXACT_FORM): 0 = FFMA2+FFMA2 (default lib), 1 = FMUL2+FFMA2, 2 = FFMA2+FADD2
mkdir -p gpurun_out
for lib in gnuradio4_b200/libgr4b200.so build/variants/libgr4b200_v1.so build/variants/libgr4b200_v2.so; do
echo "== $lib"
GR4B200_LIB=$PWD/$lib timeout 300 python scripts/time_kernels.py $((1<<28)) "fir127 exact,fir127 decim8 exact,ddc,pfb filter" 2>&1 | tee -a gpurun_out/time_fir_forms.jsonl
done
