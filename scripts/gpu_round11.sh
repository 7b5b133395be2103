#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "rotator or mathop or ddc or decimator" > gpurun_out/r11_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r11_tests.log; tail -4 gpurun_out/r11_tests.log
timeout 300 python scripts/time_kernels.py $((1<<28)) "resampler,rotator,Const,copy" > gpurun_out/time_r11.jsonl 2>&1; cat gpurun_out/time_r11.jsonl
