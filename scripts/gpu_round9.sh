#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/time_decim_tiles.jsonl
for lib in gnuradio4_b200/libgr4b200.so build/variants/libgr4b200_t128r3.so build/variants/libgr4b200_t256r3.so build/variants/libgr4b200_t64r5.so; do
echo "== $lib" | tee -a gpurun_out/time_decim_tiles.jsonl
GR4B200_LIB=$PWD/$lib timeout 300 python scripts/time_kernels.py $((1<<28)) "decim8,ddc" 2>&1 | tee -a gpurun_out/time_decim_tiles.jsonl
done
python -m pytest tests/test_cpp_host.py -q -m gpu 2>&1 | tail -3
