#!/bin/bash
# programmatic dependent launch on the FIR / FFT kernels: parity, streaming chunk sweep with and without, kernel table
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/r02v_tests.txt
cat gpurun_out/r02v_tests.txt
for pdl in 0 1 0 1; do
  echo "{\"GR4B200_PDL\": $pdl}" | tee -a gpurun_out/r02v_bm_flowgraph_pdl.jsonl
  GR4B200_PDL=$pdl timeout 300 build/cpp/bm_flowgraph --device-only --sweep --samples $((1<<29)) 2>&1 | tee -a gpurun_out/r02v_bm_flowgraph_pdl.jsonl | cut -c1-60,160-330
done
for pdl in 0 1; do GR4B200_PDL=$pdl timeout 300 python scripts/time_kernels.py $((1<<28)) "fir127 exact,fft4096,ddc,decim8 exact" 2>/dev/null | grep '"kernel"' | sed "s/^{/{\"pdl\": $pdl, /" | tee -a gpurun_out/r02v_time_kernels_pdl.jsonl | cut -c1-140; done
