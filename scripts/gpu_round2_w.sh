#!/bin/bash
# cheaper quadrant logic in the sin/cos, checkpoints of tile k+1 fetched behind the convolution of tile k, long filters on
# the CTA-wide tiles: parity, DDC / rotator / decimating FIR timing
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r02w_tests.txt
cat gpurun_out/r02w_tests.txt
for rep in 1 2; do timeout 300 python scripts/time_kernels.py $((1<<28)) "ddc,rotator,decim8 exact,fir127 exact" 2>/dev/null | grep '"kernel"' | tee -a gpurun_out/r02w_time_kernels.jsonl | cut -c1-140; done
