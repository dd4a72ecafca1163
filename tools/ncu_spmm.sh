#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:spmm_row_gather -s 5 -c 1 -o gpurun_out/prof_spmm \
    python tools/spmm_probe.py > gpurun_out/ncu_spmm.log 2>&1
ls -la gpurun_out/prof_spmm.ncu-rep
