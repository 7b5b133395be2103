#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -2
GR4B200_ROTATOR_CYCLE=0 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "rotator or ddc or Rotator" 2>&1 | tail -1
for c in 1 0; do GR4B200_ROTATOR_CYCLE=$c timeout 300 python scripts/time_mixer_checkpoints.py 2>/dev/null | sed "s/^{/{\"phase_cycle\": $c, /" | tee -a gpurun_out/r02y3_time_mixer_cycle.jsonl; done
timeout 300 python scripts/time_kernels.py $((1<<28)) "ddc,rotator" 2>/dev/null | grep '"kernel"' | cut -c1-140
