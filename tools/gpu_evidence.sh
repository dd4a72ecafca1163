#!/bin/bash
# launch list of one cfg3 step window and an ncu capture of the (time-fused) backward kernel, final build
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 260 --csv --log-file gpurun_out/launches_cfg3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
wc -l gpurun_out/launches_cfg3.csv
ncu --set full --clock-control none --import-source on -k regex:lstm16_bwd -s 7 -c 1 -o gpurun_out/prof_bwd16 \
    python tools/lstm_time.py 4096 64 12 > gpurun_out/ncu_bwd16.log 2>&1
ls -la gpurun_out/prof_bwd16.ncu-rep
