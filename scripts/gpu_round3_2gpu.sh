#!/bin/bash
# GPU box with 2 GPUs: PFB parity + timing, pipelined channelizer over NCCL, bench at N=2
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "polyphase or channelizer" > gpurun_out/pfb_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/pfb_tests.log; tail -3 gpurun_out/pfb_tests.log
timeout 300 python scripts/time_kernels.py $((1<<28)) pfb,copy > gpurun_out/time_pfb.jsonl 2>&1; cat gpurun_out/time_pfb.jsonl
for cs in 22 24 26; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/bench_pipeline.py --chunks 32 --chunk-samples $((1<<cs)) 2> gpurun_out/pipeline_2gpu_$cs.err | tee -a gpurun_out/pipeline_2gpu.jsonl
done
timeout 300 python scripts/bench_pipeline.py --chunks 32 --chunk-samples $((1<<26)) 2>/dev/null | tee -a gpurun_out/pipeline_2gpu.jsonl
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-samples $((1<<26)) > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; cat gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
