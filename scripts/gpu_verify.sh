#!/bin/bash
# Everything a round ends with, on the GPU box (1 GPU): the full parity suite, smoke(), the bench line (both arms), the
# ncu launch list of the bench command, and the kernel timing tables. Results land in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/verify_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/verify_tests.log; tail -4 gpurun_out/verify_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/verify_bench_reference.json 2> gpurun_out/verify_bench.err; cat gpurun_out/verify_bench_reference.json
python bench.py > gpurun_out/verify_bench.json 2>> gpurun_out/verify_bench.err; echo "bench exit $?"; cat gpurun_out/verify_bench.json; tail -2 gpurun_out/verify_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/verify_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-samples $((1<<24)) > gpurun_out/verify_bench_under_ncu.log 2>&1
timeout 600 python scripts/time_kernels.py $((1<<28)) > gpurun_out/verify_time_kernels.jsonl 2>&1; cat gpurun_out/verify_time_kernels.jsonl
timeout 300 python scripts/time_fft.py $((1<<28)) 2>&1 | grep -v "direct loads" > gpurun_out/verify_time_fft.jsonl
