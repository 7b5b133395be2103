#!/bin/bash
# GPU box: FFT parity tests, A/B timing, ncu captures of the 4096 kernels (spectrum + block)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fft or flowgraph or channelizer" > gpurun_out/fft_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/fft_tests.log
tail -5 gpurun_out/fft_tests.log
timeout 600 python scripts/time_fft.py $((1<<28)) > gpurun_out/time_fft.jsonl 2> gpurun_out/time_fft.err
cat gpurun_out/time_fft.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fftRadix -c 2 -s 2 -f -o gpurun_out/prof_fft_radix python scripts/profile_kernels.py fftc2c > gpurun_out/ncu_c2c.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fftRadix -c 1 -s 2 -f -o gpurun_out/prof_fft_radix_block python scripts/profile_kernels.py fftblock > gpurun_out/ncu_block.log 2>&1
GR4B200_FFT_TMA=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:fftRadix -c 1 -s 2 -f -o gpurun_out/prof_fft_radix_direct python scripts/profile_kernels.py fftc2c > gpurun_out/ncu_c2c_direct.log 2>&1
ls -la gpurun_out
