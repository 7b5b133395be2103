#!/bin/bash
# Round 2, session 2, first measurement: taps from the kernel parameters (uniform registers) vs shared-memory tables, /8 tile
# variants on top, fused DDC with the carried halo (grid waves), full-rate FIR at one CTA per SM, scheduler host profile.
mkdir -p gpurun_out
O=gpurun_out/r02m_time_variants.jsonl
: > $O
python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py -q -m gpu -x -k "fir or ddc or rotator or mixer or Rotator or golden" 2>&1 | tail -3 > gpurun_out/r02m_tests.txt
cat gpurun_out/r02m_tests.txt
run() { # label, env..., -- kernels
  label=$1; shift
  env "$@" timeout 300 python scripts/time_kernels.py $((1<<28)) "$KERNELS" 2>/dev/null | grep '"kernel"' | sed "s/^{/{\"cfg\": \"$label\", /" >> $O
}
KERNELS="fir127 exact,fir127 fast,decim8,ddc,fir127 decim2 exact,fir127 decim4 exact,fir127 decim16 exact,rotator"
run "param_taps=0" GR4B200_FIR_PARAM_TAPS=0
run "param_taps=1" GR4B200_FIR_PARAM_TAPS=1
KERNELS="decim8 exact,ddc"
for v in 0 2 4 5 8 9; do run "variant=$v" GR4B200_DECIM8_VARIANT=$v; done
for v in 0 2 4 8; do run "variant=$v param_taps=0" GR4B200_DECIM8_VARIANT=$v GR4B200_FIR_PARAM_TAPS=0; done
KERNELS="ddc"
for m in 1 2 3 4 8 16; do run "ddc grid_mult=$m" GR4B200_FIR_GRID_MULT=$m; done
for v in 4 8; do for m in 2 4; do run "ddc variant=$v grid_mult=$m" GR4B200_DECIM8_VARIANT=$v GR4B200_FIR_GRID_MULT=$m; done; done
KERNELS="fir127 exact"
run "full rate, one CTA per SM" GR4B200_FIR_EXTRA_SMEM=50000
run "full rate, one CTA per SM, param_taps=0" GR4B200_FIR_EXTRA_SMEM=50000 GR4B200_FIR_PARAM_TAPS=0
cat $O
for c in 65536 1048576; do
  GR4B200_SCHED_PROFILE=1 timeout 120 build/cpp/bm_flowgraph --device-only --chunk $c --samples $((1<<28)) > gpurun_out/r02m_sched_profile_$c.txt 2>&1
  tail -12 gpurun_out/r02m_sched_profile_$c.txt
done
