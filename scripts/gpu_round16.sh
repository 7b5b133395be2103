#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/time_grid_variants2.jsonl
run() { echo "== $1" | tee -a gpurun_out/time_grid_variants2.jsonl; env $1 timeout 300 python scripts/time_kernels.py $((1<<28)) "$2" 2>&1 | tee -a gpurun_out/time_grid_variants2.jsonl; }
run "X=default" "fft4096,fft256,fft1024,pfb,resampler 160,resampler 1/1,copy"
run "GR4B200_PFB_CTAS=16" "pfb filter,pfb channelizer, fused"
run "GR4B200_PFB_CTAS=32" "pfb filter,pfb channelizer, fused"
run "GR4B200_PFB_CTAS=4" "pfb filter,pfb channelizer, fused"
run "GR4B200_RESAMPLER_GRID_MULT=4" "resampler 160,resampler 1/1"
run "GR4B200_RESAMPLER_GRID_MULT=0" "resampler 160,resampler 1/1"
timeout 300 python scripts/time_fft.py $((1<<28)) 2>&1 | grep -v "direct loads" > gpurun_out/time_fft_r16.jsonl; cat gpurun_out/time_fft_r16.jsonl
