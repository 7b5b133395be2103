#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "resampler or polyphase or channelizer" > gpurun_out/resampler_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/resampler_tests.log; tail -12 gpurun_out/resampler_tests.log
timeout 300 python scripts/time_kernels.py $((1<<28)) "resampler,pfb,copy" > gpurun_out/time_resampler.jsonl 2>&1; cat gpurun_out/time_resampler.jsonl
