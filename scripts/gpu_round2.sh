#!/bin/bash
# GPU box (1 GPU): full parity suite, bench line, launch list of the bench command, kernel timings, ncu of the PFB kernel
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests.log
tail -4 gpurun_out/gpu_tests.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-samples $((1<<24)) > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 python scripts/time_kernels.py $((1<<28)) > gpurun_out/time_kernels.jsonl 2> gpurun_out/time_kernels.err; cat gpurun_out/time_kernels.jsonl
timeout 300 python scripts/time_fft.py $((1<<28)) > gpurun_out/time_fft2.jsonl 2>&1; grep -E "fft(4096|1024|8192|256) " gpurun_out/time_fft2.jsonl | grep -v legacy
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pfbStream -c 1 -s 2 -f -o gpurun_out/prof_pfb python scripts/profile_kernels.py pfb > gpurun_out/ncu_pfb.log 2>&1
timeout 300 python scripts/bench_pipeline.py --chunks 16 > gpurun_out/pipeline_1gpu.json 2> gpurun_out/pipeline_1gpu.err; cat gpurun_out/pipeline_1gpu.json
ls -la gpurun_out | head -40
