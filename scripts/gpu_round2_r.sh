#!/bin/bash
# tap-mode defaults, /4 and /16 tile variants, quarter-size tiles for short full-rate calls (streaming chunk sweep)
mkdir -p gpurun_out
O=gpurun_out/r02r_time_variants.jsonl
: > $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r02r_tests.txt
cat gpurun_out/r02r_tests.txt
run() { label=$1; shift; env "$@" timeout 300 python scripts/time_kernels.py $((1<<28)) "$KERNELS" 2>/dev/null | grep '"kernel"' | sed "s/^{/{\"cfg\": \"$label\", /" >> $O; }
KERNELS="fir127 exact,fir127 fast,decim8,ddc,fir127 decim2 exact,fir127 decim4 exact,fir127 decim16 exact"
run "defaults" GR4B200_NOP=1
KERNELS="decim4 exact"
for v in 1 2 3; do run "decim4 variant=$v" GR4B200_DECIM4_VARIANT=$v; run "decim4 variant=$v tap_mode=2" GR4B200_DECIM4_VARIANT=$v GR4B200_FIR_TAP_MODE=2; done
KERNELS="decim16 exact"
for v in 1 2 3; do run "decim16 variant=$v" GR4B200_DECIM16_VARIANT=$v;  run "decim16 variant=$v tap_mode=0" GR4B200_DECIM16_VARIANT=$v GR4B200_FIR_TAP_MODE=0; done
cat $O | cut -c1-150
timeout 300 build/cpp/bm_flowgraph --device-only --sweep --samples $((1<<29)) 2>&1 | cut -c1-60,160-330 | tee gpurun_out/r02r_bm_flowgraph_sweep.txt
