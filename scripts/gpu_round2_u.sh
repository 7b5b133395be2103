#!/bin/bash
# FIR of chunk k+1 next to the FFT block of chunk k on two streams: FIR at one CTA per SM (extra shared memory) so that an
# FFT CTA fits beside it, one resident wave or sixteen, FFT stream at high priority or not
mkdir -p gpurun_out
O=gpurun_out/r02u_time_overlap.jsonl
for extra in 0 50000; do for mult in 16 1; do for prio in 0 1; do
  echo "{\"fir_extra_smem\": $extra, \"fir_grid_mult\": $mult, \"fft_stream_high_priority\": $prio}" >> $O
  OVERLAP_CHUNKS=8,32 OVERLAP_PRIORITY=$prio GR4B200_FIR_GRID_MULT=$mult GR4B200_FIR_EXTRA_SMEM=$extra timeout 300 python scripts/time_overlap.py $((1<<29)) 2>&1 | grep two_streams >> $O
done; done; done
cat $O
