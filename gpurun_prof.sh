mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lstm_cell_tc --launch-skip 112 --launch-count 2 -o gpurun_out/r1_fwd_cell_r6 -f $B > gpurun_out/ncu_fwd.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lstm_bwd_tc --launch-skip 111 --launch-count 2 -o gpurun_out/r1_bwd_r5 -f $B > gpurun_out/ncu_bwd.log 2>&1
tail -3 gpurun_out/ncu_fwd.log gpurun_out/ncu_bwd.log
