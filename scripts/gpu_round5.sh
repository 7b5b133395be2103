#!/bin/bash
# GPU box (1 GPU): fused FIR->FFT parity + timing + ncu
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fused_fir_fft or flowgraph" > gpurun_out/firfft_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/firfft_tests.log; tail -15 gpurun_out/firfft_tests.log
timeout 300 python scripts/time_kernels.py $((1<<28)) "fir127 exact,fir127 fast,fft4096 block,copy" > gpurun_out/time_firfft.jsonl 2>&1; cat gpurun_out/time_firfft.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:firFftBlock -c 1 -s 2 -f -o gpurun_out/prof_firfft python scripts/profile_kernels.py firfft > gpurun_out/ncu_firfft.log 2>&1; tail -2 gpurun_out/ncu_firfft.log
