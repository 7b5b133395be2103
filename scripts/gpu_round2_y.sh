#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
timeout 300 python scripts/time_kernels.py $((1<<28)) "fir127,ddc,rotator" 2>/dev/null | grep '"kernel"' | cut -c1-130
