#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -2
timeout 300 python scripts/time_mixer_checkpoints.py 2>/dev/null | tee gpurun_out/r02y2_time_mixer.jsonl
cap() {
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -s $4 -f -o /tmp/r02y2_$1 python scripts/profile_kernels.py $3 > gpurun_out/r02y2_$1.log 2>&1; echo "$1: exit $?"
  python scripts/summarize_ncu.py /tmp/r02y2_$1.ncu-rep gpurun_out/r02y2_$1_ncu.md "$5"
  rm -f /tmp/r02y2_$1.ncu-rep
}
cap checkpoint checkpointKernel ddc 2 "checkpointKernel<512> of a fused DDC call, 2^26 samples"
cap rotate rotateKernel rot 2 "rotateKernel<+1> (in-range fast path), 2^26 samples"
