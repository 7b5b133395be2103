#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/time_mathop_variants.jsonl
for lib in gnuradio4_b200/libgr4b200.so build/variants/libgr4b200_m1.so build/variants/libgr4b200_m2.so build/variants/libgr4b200_m3.so build/variants/libgr4b200_m4.so build/variants/libgr4b200_m5.so; do
echo "== $lib" | tee -a gpurun_out/time_mathop_variants.jsonl
GR4B200_LIB=$PWD/$lib timeout 300 python scripts/time_kernels.py $((1<<28)) "Const,copy" 2>&1 | tee -a gpurun_out/time_mathop_variants.jsonl
done
