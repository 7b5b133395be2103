#!/bin/bash
# Everything a round ends with, on the GPU box (1 GPU): the full parity suite, smoke(), the bench line (both arms), the
# ncu launch list of the bench command, the kernel timing tables, the C++ flowgraph benchmark and the host-link ceiling.
# Results land in gpurun_out/ (copy what should be judged into profiles/).
P=${1:-verify}
mkdir -p gpurun_out; rm -f gpurun_out/${P}_time_tap_modes.jsonl
python -m pytest tests -q -m gpu > gpurun_out/${P}_gpu_tests.txt 2>&1; echo "tests exit $?" >> gpurun_out/${P}_gpu_tests.txt; grep -v "^Exception ignored\|^Traceback (most\|blocks.py\|AttributeError" gpurun_out/${P}_gpu_tests.txt | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_reference.json 2> gpurun_out/${P}_bench.err; cut -c1-300 gpurun_out/${P}_bench_reference.json
python bench.py --steps 20 --warmup 5 > gpurun_out/${P}_bench.json 2>> gpurun_out/${P}_bench.err; echo "bench exit $?"; cut -c1-400 gpurun_out/${P}_bench.json; tail -2 gpurun_out/${P}_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${P}_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-samples $((1<<24)) > gpurun_out/${P}_bench_under_ncu.log 2>&1
timeout 600 python scripts/time_kernels.py $((1<<28)) 2>/dev/null > gpurun_out/${P}_time_kernels.jsonl; cat gpurun_out/${P}_time_kernels.jsonl | cut -c1-160
timeout 300 python scripts/time_fft.py $((1<<28)) 2>&1 | grep -v "direct loads" > gpurun_out/${P}_time_fft.jsonl
build/cpp/bm_flowgraph --sweep --samples $((1<<29)) > gpurun_out/${P}_bm_flowgraph.jsonl 2>&1
for v in 1 2 3; do build/cpp/bm_flowgraph --host-only --variant $v --samples $((1<<28)) >> gpurun_out/${P}_bm_flowgraph.jsonl 2>&1; done
python scripts/time_host_link.py > gpurun_out/${P}_host_link_1gpu.json 2>/dev/null
cuobjdump -sass gnuradio4_b200/libgr4b200.so 2>/dev/null | grep -o "^\s*/\*[0-9a-f]*\*/\s*[A-Z0-9_.]*" | awk '{print $2}' | sed 's/\..*//' | sort | uniq -c | sort -rn > gpurun_out/${P}_sass_opcodes.txt
for m in 0 1 2; do GR4B200_FIR_TAP_MODE=$m timeout 300 python scripts/time_kernels.py $((1<<28)) "ddc,fir127 exact,decim8 exact,decim4 exact" 2>/dev/null | grep '"kernel"' | sed "s/^{/{\"tap_mode\": $m, /" >> gpurun_out/${P}_time_tap_modes.jsonl; done
