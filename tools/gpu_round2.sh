#!/bin/bash
# final single-GPU evidence: cfg3 bench, launch list, small-config step times with and without CUDA graphs, memcheck
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3_n1.json 2> gpurun_out/bench_cfg3_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_cfg3_n1.json"))
print("cfg3", round(d["ms_per_step"], 2), round(d["value"] / 1e6, 1), round(d["e2e"]["value"] / 1e6, 1), d["loss_check"]["ok"], d["loss_check"]["rel_err"],
      round(d["roofline"]["frac"], 3), d["roofline"]["traffic"], round(d["roofline_lstm"]["forward"]["ms"], 2), round(d["roofline_lstm"]["backward"]["ms"], 2),
      d["gpu_launches"], d["clocks"], round(d["cpu_baseline"]["value"]))
PY
for wl in cfg1 cfg2; do for g in "" "--cuda-graph"; do
  python bench.py --workload $wl --arith fp32 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e $g 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$wl', '$g', 'ms/step', round(d['ms_per_step'],3), 'M r-t/s', round(d['value']/1e6,2), d['loss_check']['ok'], d['gpu_launches'])"
done; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 400 --csv --log-file gpurun_out/launches_cfg3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
wc -l gpurun_out/launches_cfg3.csv
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -q -x -k "tensor_core or golden or bf16_arith" > gpurun_out/sanitizer_memcheck.log 2>&1
tail -4 gpurun_out/sanitizer_memcheck.log
