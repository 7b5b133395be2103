#!/bin/bash
# defaults after the tile / tap-source changes; /2 variants; DDC repeatability (default vs GR4B200_FIR_GRID_MULT)
mkdir -p gpurun_out
O=gpurun_out/r02s_time_variants.jsonl
: > $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r02s_tests.txt
cat gpurun_out/r02s_tests.txt
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv,noheader
run() { label=$1; shift; env "$@" timeout 300 python scripts/time_kernels.py $((1<<28)) "$KERNELS" 2>/dev/null | grep '"kernel"' | sed "s/^{/{\"cfg\": \"$label\", /" >> $O; }
KERNELS="fir127 exact,fir127 fast,decim8,ddc,fir127 decim2 exact,fir127 decim4 exact,fir127 decim16 exact,rotator,copy,fft4096 block"
run "defaults" GR4B200_NOP=1
run "defaults again" GR4B200_NOP=1
KERNELS="ddc,rotator"
for g in 2 16; do run "ddc grid_mult=$g" GR4B200_FIR_GRID_MULT=$g; done
run "ddc tap_mode=1" GR4B200_FIR_TAP_MODE=1
run "ddc variant=2" GR4B200_DECIM8_VARIANT=2
KERNELS="decim2 exact"
for v in 1 2; do run "decim2 variant=$v" GR4B200_DECIM2_VARIANT=$v; run "decim2 variant=$v tap_mode=0" GR4B200_DECIM2_VARIANT=$v GR4B200_FIR_TAP_MODE=0; done
KERNELS="decim4 exact"
for v in 0 2 4; do run "decim4 variant=$v" GR4B200_DECIM4_VARIANT=$v; done
KERNELS="decim16 exact"
for v in 0 1 4; do run "decim16 variant=$v" GR4B200_DECIM16_VARIANT=$v; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv,noheader
cat $O | cut -c1-150
