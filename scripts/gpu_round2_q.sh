#!/bin/bash
# new defaults (warp-sized single-stage /8 tiles with parameter taps, also for the fused DDC): parity, timing, source profile
mkdir -p gpurun_out
O=gpurun_out/r02q_time_variants.jsonl
: > $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_fullsize.py -q -m gpu -x -k "fir or ddc or rotator or mixer or Rotator or golden or DDC" 2>&1 | tail -3 > gpurun_out/r02q_tests.txt
cat gpurun_out/r02q_tests.txt
run() { label=$1; shift; env "$@" timeout 300 python scripts/time_kernels.py $((1<<28)) "$KERNELS" 2>/dev/null | grep '"kernel"' | sed "s/^{/{\"cfg\": \"$label\", /" >> $O; }
KERNELS="fir127 exact,fir127 fast,decim8,ddc,fir127 decim2 exact,fir127 decim4 exact,fir127 decim16 exact,rotator"
run "defaults" GR4B200_NOP=1
KERNELS="decim8,ddc"
for m in 0 1 2; do for v in 0 4 12 13 16; do run "variant=$v tap_mode=$m" GR4B200_DECIM8_VARIANT=$v GR4B200_FIR_TAP_MODE=$m; done; done
for g in 1 2 4 8 16; do run "default grid_mult=$g" GR4B200_FIR_GRID_MULT=$g; done
cat $O | cut -c1-170
sed -i 's/r02o_/r02q_/g' scripts/gpu_round2_o.sh
bash scripts/gpu_round2_o.sh > /dev/null 2>&1
ls gpurun_out
