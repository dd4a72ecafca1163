#!/bin/bash
# validation of the step-grouped time-fused backward: quick parity subset, the cfg5-shape strict test, kernel timing, cfg3 bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -s -k "lstm or tensor_core or golden or cfg3 or bf16 or training or (cfg5 and False)" 2>&1 | grep -E "passed|failed|AssertionError|max-norm relative errors" | cut -c1-600 | tail -8
python tools/lstm_time.py 2>&1 | tail -1
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3_n1.json 2> gpurun_out/bench_cfg3_n1.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_cfg3_n1.json"))
print("cfg3", round(d["ms_per_step"], 2), round(d["value"] / 1e6, 1), round(d["e2e"]["value"] / 1e6, 1), d["loss_check"]["ok"], d["loss_check"]["rel_err"],
      round(d["roofline"]["frac"], 3), round(d["roofline_lstm"]["forward"]["ms"], 2), round(d["roofline_lstm"]["backward"]["ms"], 2), d["gpu_launches"], d["clocks"])
PY
