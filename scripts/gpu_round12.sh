#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r12_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r12_tests.log; tail -4 gpurun_out/r12_tests.log
timeout 300 python scripts/time_kernels.py $((1<<28)) > gpurun_out/time_r12.jsonl 2>&1; cat gpurun_out/time_r12.jsonl
python bench.py --workload ddc_fft --steps 3 --warmup 3 > gpurun_out/bench_ddc_r12.json 2>/dev/null; cat gpurun_out/bench_ddc_r12.json
