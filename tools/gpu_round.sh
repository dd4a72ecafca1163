#!/bin/bash
# One GPU-box round: full GPU test-suite, ncu captures, bench lines (cfg3 fp32-grade, cfg2 bf16 / fp32-grade).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300 > gpurun_out/tests_gpu.log; cat gpurun_out/tests_gpu.log
timeout 600 ./tools/ncu_lstm16.sh 2>&1 | tail -2
timeout 300 ./tools/ncu_spmm.sh 2>&1 | tail -1
python tools/spmm_probe.py
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3_n1.json 2> gpurun_out/bench_cfg3_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_cfg3_n1.json"))
print("cfg3", d["ms_per_step"], d["value"], d["e2e"]["value"], d["loss_check"], d["roofline"]["frac"],
      d["roofline_lstm"]["forward"]["ms"], d["roofline_lstm"]["backward"]["ms"], d["clocks"])
PY
tail -2 gpurun_out/bench_cfg3_n1.err
python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2_bf16_n1.json 2> gpurun_out/bench_cfg2.err
python bench.py --workload cfg2 --arith fp32 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2_fp32_n1.json 2>> gpurun_out/bench_cfg2.err
python - <<'PY'
import json
for f in ("bench_cfg2_bf16_n1", "bench_cfg2_fp32_n1"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, d["ms_per_step"], d["value"], d["e2e"]["value"], d["loss_check"], d["dtype"], d["gpu_launches"])
PY
