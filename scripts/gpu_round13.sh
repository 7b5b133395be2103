#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_cpp_host.py -q -m gpu -k "rotator or mathop or ddc or qa_device" > gpurun_out/r13_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/r13_tests.log; tail -4 gpurun_out/r13_tests.log
timeout 300 python scripts/time_kernels.py $((1<<28)) "rotator,Const,ddc,copy" > gpurun_out/time_r13.jsonl 2>&1; cat gpurun_out/time_r13.jsonl
