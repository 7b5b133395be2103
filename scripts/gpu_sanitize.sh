#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
GR4B200_FIR_GRID_MULT=1 timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_kernels.py > gpurun_out/sanitize_$tool.log 2>&1
echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize run done|Error|hazard" gpurun_out/sanitize_$tool.log | head -12
done
