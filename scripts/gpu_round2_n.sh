#!/bin/bash
# Tap source A/B/A/B (0 shared scalars, 1 shared pairs, 2 parameter pairs), /8 tile variants per tap source, DDC waves,
# ncu --set full of the /8 FIR and the fused DDC as they are now.
mkdir -p gpurun_out
O=gpurun_out/r02n_time_variants.jsonl
: > $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -m gpu -x -k "fir or ddc or rotator or mixer or Rotator or golden" 2>&1 | tail -3 > gpurun_out/r02n_tests.txt
cat gpurun_out/r02n_tests.txt
run() { label=$1; shift; env "$@" timeout 300 python scripts/time_kernels.py $((1<<28)) "$KERNELS" 2>/dev/null | grep '"kernel"' | sed "s/^{/{\"cfg\": \"$label\", /" >> $O; }
KERNELS="fir127 exact,fir127 fast,decim8,ddc,fir127 decim2 exact,fir127 decim4 exact,fir127 decim16 exact"
for rep in 1 2; do for m in 0 1 2; do run "tap_mode=$m rep=$rep" GR4B200_FIR_TAP_MODE=$m; done; done
KERNELS="decim8 exact,ddc"
for m in 1 2; do for v in 0 2 4 5 8; do run "variant=$v tap_mode=$m" GR4B200_DECIM8_VARIANT=$v GR4B200_FIR_TAP_MODE=$m; done; done
KERNELS="ddc"
for g in 2 8 16 32 0; do run "ddc grid_mult=$g" GR4B200_FIR_GRID_MULT=$g; done
for v in 4 8; do for g in 8 16; do run "ddc variant=$v grid_mult=$g" GR4B200_DECIM8_VARIANT=$v GR4B200_FIR_GRID_MULT=$g; done; done
cat $O
cap() { timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -s $4 -f -o gpurun_out/r02n_$1 python scripts/profile_kernels.py $3 > gpurun_out/r02n_$1.log 2>&1; echo "$1: exit $?"; }
cap ddc firDecim ddc 2
cap firdecim firDecim firdecim 2
GR4B200_DECIM8_VARIANT=4 cap firdecim_v4 firDecim firdecim 2
ls -la gpurun_out/r02n_*.ncu-rep
