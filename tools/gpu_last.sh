#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3_n1.json 2> gpurun_out/bench_cfg3_n1.err
python bench.py --workload cfg2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2_bf16_n1.json 2> gpurun_out/bench_cfg2.err
python bench.py --workload cfg2 --arith fp32 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_cfg2_fp32_n1.json 2>> gpurun_out/bench_cfg2.err
python - <<'PY'
import json
for f in ("bench_cfg3_n1", "bench_cfg2_bf16_n1", "bench_cfg2_fp32_n1"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, round(d["ms_per_step"], 3), round(d["value"] / 1e6, 1), round(d["e2e"]["value"] / 1e6, 1), d["loss_check"]["ok"], d["gpu_launches"], d["clocks"]["sm_mhz"],
          round(d["roofline_lstm"]["forward"]["ms"], 2) if "roofline_lstm" in d else "", round(d["roofline_lstm"]["backward"]["ms"], 2) if "roofline_lstm" in d else "")
PY
