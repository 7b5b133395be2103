#!/bin/bash
# /8 decimating FIR and fused DDC: tile shape x stage count variants (GR4B200_DECIM8_VARIANT), 2^28 input samples
mkdir -p gpurun_out
: > gpurun_out/r02c_time_decim8_variants.jsonl
for v in -1 0 1 2 3 4 5 6 7; do
  GR4B200_DECIM8_VARIANT=$v timeout 300 python scripts/time_kernels.py $((1<<28)) "decim8,ddc" 2>/dev/null | sed "s/^{/{\"variant\": $v, /" >> gpurun_out/r02c_time_decim8_variants.jsonl
done
cat gpurun_out/r02c_time_decim8_variants.jsonl
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "ddc or fir" 2>&1 | tail -2
