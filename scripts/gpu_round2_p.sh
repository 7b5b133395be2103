#!/bin/bash
# /8 FIR and fused DDC: warp-sized CTAs (variants 10-15) next to the CTA-wide tiles, per tap source; lane-offset table
mkdir -p gpurun_out
O=gpurun_out/r02p_time_variants.jsonl
: > $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -m gpu -x -k "fir or ddc or rotator or mixer or Rotator or golden" 2>&1 | tail -3 > gpurun_out/r02p_tests.txt
cat gpurun_out/r02p_tests.txt
run() { label=$1; shift; env "$@" timeout 300 python scripts/time_kernels.py $((1<<28)) "$KERNELS" 2>/dev/null | grep '"kernel"' | sed "s/^{/{\"cfg\": \"$label\", /" >> $O; }
KERNELS="fir127 exact,fir127 fast,decim8,ddc,fir127 decim2 exact,fir127 decim4 exact,fir127 decim16 exact"
run "defaults" GR4B200_NOP=1
KERNELS="decim8 exact,ddc"
for m in 1 2; do for v in 0 2 4 8 10 11 12 13 14 15; do run "variant=$v tap_mode=$m" GR4B200_DECIM8_VARIANT=$v GR4B200_FIR_TAP_MODE=$m; done; done
for v in 12 13; do for g in 2 4 16; do run "variant=$v tap_mode=1 grid_mult=$g" GR4B200_DECIM8_VARIANT=$v GR4B200_FIR_TAP_MODE=1 GR4B200_FIR_GRID_MULT=$g; done; done
cat $O | cut -c1-170
