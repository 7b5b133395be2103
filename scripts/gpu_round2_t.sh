#!/bin/bash
# defaults after the pointer-walk fix (uniform tap loads in every kernel), all parity tests, streaming chunk sweep
mkdir -p gpurun_out
O=gpurun_out/r02t_time_kernels.jsonl
python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r02t_tests.txt
cat gpurun_out/r02t_tests.txt
timeout 600 python scripts/time_kernels.py $((1<<28)) 2>/dev/null > $O
cut -c1-120 $O
GR4B200_DECIM2_VARIANT=0 GR4B200_DECIM4_VARIANT=0 GR4B200_DECIM16_VARIANT=0 GR4B200_DECIM8_VARIANT=0 GR4B200_FIR_TAP_MODE=0 timeout 300 python scripts/time_kernels.py $((1<<28)) "fir127,ddc" 2>/dev/null | grep '"kernel"' | sed "s/^{/{\"cfg\": \"round-1 tiles and scalar shared taps\", /" > gpurun_out/r02t_time_old_tiles.jsonl
cut -c1-150 gpurun_out/r02t_time_old_tiles.jsonl
timeout 300 build/cpp/bm_flowgraph --device-only --sweep --samples $((1<<29)) 2>&1 | tee gpurun_out/r02t_bm_flowgraph_sweep.jsonl | cut -c1-60,160-330
