#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/gpu_tests8.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests8.log; tail -6 gpurun_out/gpu_tests8.log
rm -f gpurun_out/pipeline8.jsonl
for n in 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n scripts/bench_pipeline.py --chunks 32 --chunk-samples $((1<<26)) 2> gpurun_out/pipeline8_${n}gpu.err | tee -a gpurun_out/pipeline8.jsonl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2958$n scripts/bench_pipeline.py --chunks 64 --chunk-samples $((1<<24)) 2>> gpurun_out/pipeline8_${n}gpu.err | tee -a gpurun_out/pipeline8.jsonl
done
tail -3 gpurun_out/pipeline8_4gpu.err
