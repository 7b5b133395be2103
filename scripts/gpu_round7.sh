#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests7.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests7.log; tail -6 gpurun_out/gpu_tests7.log
python bench.py > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "bench exit $?"; cat gpurun_out/bench7.json; tail -3 gpurun_out/bench7.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches7.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-samples $((1<<24)) > gpurun_out/bench_under_ncu7.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:firKernel -c 1 -s 2 -f -o gpurun_out/prof_fir7 python scripts/profile_kernels.py fir > gpurun_out/ncu_fir7.log 2>&1
