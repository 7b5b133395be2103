#!/bin/bash
# ncu --set full captures of the mixer kernels with the library-exact sin/cos (FP64 pipe): fused DDC, rotator
mkdir -p gpurun_out
cap() { timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -c 1 -s $4 -f -o gpurun_out/r02g_$1 python scripts/profile_kernels.py $3 > gpurun_out/r02g_$1.log 2>&1; echo "$1: exit $?"; }
cap ddc firDecim ddc 2
cap rotator rotateKernel rot 2
ls -la gpurun_out/r02g_*.ncu-rep
