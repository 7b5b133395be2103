#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -2
timeout 300 python scripts/time_mixer_checkpoints.py 2>/dev/null | tee -a gpurun_out/r02x_time_mixer_checkpoints.jsonl
