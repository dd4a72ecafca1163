#!/bin/bash
# ncu captures of the layer-0 variants of the LSTM kernels at cfg3 size (one steady-state launch each)
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:lstm16_fwd_kernel<\(int\)2, \(int\)1>' -s 14 -c 1 -o gpurun_out/prof_fwd16_l0 \
    python tools/lstm_time.py 4096 64 12 > gpurun_out/ncu_fwd16_l0.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:lstm16_bwd_kernel<\(int\)2, \(int\)1>' -s 14 -c 1 -o gpurun_out/prof_bwd16_l0 \
    python tools/lstm_time.py 4096 64 12 > gpurun_out/ncu_bwd16_l0.log 2>&1
ls -la gpurun_out/*_l0.ncu-rep
