#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -2
for m in 2 1 0 2; do GR4B200_FIR_TAP_MODE=$m timeout 300 python scripts/time_kernels.py $((1<<28)) "ddc,fir127 exact,decim8 exact,rotator" 2>/dev/null | grep '"kernel"' | sed "s/^{/{\"tap_mode\": $m, /" | cut -c1-150; done
timeout 300 python scripts/time_kernels.py $((1<<28)) "fir127,ddc,pfb" 2>/dev/null | grep '"kernel"' | cut -c1-130
