#!/bin/bash
# ncu captures of the second-generation LSTM kernels at cfg3 size (one steady-state launch each: layer 1, t = 1 / T-2).
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lstm16_fwd -s 40 -c 1 -o gpurun_out/prof_fwd16 \
    python tools/lstm_time.py 4096 64 12 > gpurun_out/ncu_fwd16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lstm16_bwd -s 40 -c 1 -o gpurun_out/prof_bwd16 \
    python tools/lstm_time.py 4096 64 12 > gpurun_out/ncu_bwd16.log 2>&1
ls -la gpurun_out/*.ncu-rep
