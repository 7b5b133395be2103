#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/time_grid_variants.jsonl
run() { echo "== $1" | tee -a gpurun_out/time_grid_variants.jsonl; env $1 timeout 300 python scripts/time_kernels.py $((1<<28)) "$2" 2>&1 | tee -a gpurun_out/time_grid_variants.jsonl; }
run "X=default" "rotator,Const,fft4096,fft256,fft1024,pfb fft,copy"
run "GR4B200_ROTATOR_CTAS=8" "rotator"
run "GR4B200_FFT_GRID_MULT=0" "fft4096,fft256,fft1024"
run "GR4B200_FFT_GRID_MULT=4" "fft4096,fft256,fft1024"
run "GR4B200_FFT_GRID_MULT=16" "fft4096,fft256,fft1024"
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rotator or mathop or fft" 2>&1 | tail -3
