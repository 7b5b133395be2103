#!/bin/bash
# one `ncu --set full` capture per kernel family with the current code (2^26 samples each); summaries go to profiles/
mkdir -p gpurun_out
cap() { # name, kernel regex, profile_kernels.py mode, launch-skip
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -s $4 -f -o gpurun_out/final_$1 python scripts/profile_kernels.py $3 > gpurun_out/final_$1.log 2>&1
  echo "$1: exit $?"
}
cap fft_c2c fftRadix fftc2c 2
cap fft_block fftRadix fftblock 2
cap fir_decim8 firDecim firdecim 2
cap ddc firDecim ddc 2
cap rotator rotateKernel rot 2
cap mathop mathopConst math 2
cap pfb pfbStream pfb 2
cap channelizer pfbChannelizer channelizer 2
cap resampler resamplerKernel resampler 2
cap fir_fast firKernel firfast 2
ls -la gpurun_out/final_*.ncu-rep
