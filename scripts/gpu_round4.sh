#!/bin/bash
# GPU box (4 GPUs): full parity suite incl. full-size tests on GPU 0, PFB/channelizer timings, ncu of the fused channelizer,
# pipelined channelizer on 2 and 4 GPUs, DDC bench workload on 4 GPUs
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests4.log 2>&1; echo "tests exit $?" >> gpurun_out/gpu_tests4.log; tail -5 gpurun_out/gpu_tests4.log
timeout 300 python scripts/time_kernels.py $((1<<28)) pfb,copy,ddc,rotator > gpurun_out/time_pfb4.jsonl 2>&1; cat gpurun_out/time_pfb4.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pfbChannelizer -c 1 -s 2 -f -o gpurun_out/prof_channelizer python scripts/profile_kernels.py channelizer > gpurun_out/ncu_channelizer.log 2>&1
rm -f gpurun_out/pipeline_ngpu.jsonl
for n in 2 4; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n scripts/bench_pipeline.py --chunks 32 --chunk-samples $((1<<26)) 2> gpurun_out/pipeline_${n}gpu.err | tee -a gpurun_out/pipeline_ngpu.jsonl
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n scripts/bench_pipeline.py --chunks 64 --chunk-samples $((1<<24)) 2>> gpurun_out/pipeline_${n}gpu.err | tee -a gpurun_out/pipeline_ngpu.jsonl
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 3 --warmup 3 --workload ddc_fft > gpurun_out/bench_ddc_4gpu.json 2> gpurun_out/bench_ddc_4gpu.err; cat gpurun_out/bench_ddc_4gpu.json; tail -2 gpurun_out/bench_ddc_4gpu.err
timeout 600 python bench.py --steps 3 --warmup 3 --workload ddc_fft > gpurun_out/bench_ddc_1gpu.json 2>> gpurun_out/bench_ddc_4gpu.err; cat gpurun_out/bench_ddc_1gpu.json
