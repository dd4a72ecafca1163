mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r9_n1.json 2> gpurun_out/bench_r9_n1.err; tail -c 400 gpurun_out/bench_r9_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/r1_launches_cfg3_r9.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/launch_r9.log 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lstm_cell_tc --launch-skip 112 --launch-count 1 -o gpurun_out/r1_fwd_cell_r9 -f $B > gpurun_out/ncu_fwd.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:lstm_bwd_tc --launch-skip 111 --launch-count 2 -o gpurun_out/r1_bwd_r9 -f $B > gpurun_out/ncu_bwd.log 2>&1
ls -la gpurun_out/*r9*
