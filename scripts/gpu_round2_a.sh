#!/bin/bash
# round 2, first GPU call: parity suite with the bit-exact mixer, kernel timings of the mixer / DDC, host link ceiling, bench
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x > gpurun_out/r02a_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r02a_tests.log; tail -15 gpurun_out/r02a_tests.log
timeout 600 python scripts/time_kernels.py $((1<<28)) "rotator,ddc,MultiplyConst,fir127 decim8" > gpurun_out/r02a_time_kernels.jsonl 2>&1; cat gpurun_out/r02a_time_kernels.jsonl
timeout 300 python scripts/time_host_link.py > gpurun_out/r02a_host_link_1gpu.json 2>&1; cat gpurun_out/r02a_host_link_1gpu.json
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; echo "bench exit $?"; cat gpurun_out/r02a_bench.json; tail -3 gpurun_out/r02a_bench.err
nvidia-smi topo -m > gpurun_out/r02a_topo.txt 2>&1; lscpu | head -25 >> gpurun_out/r02a_topo.txt
