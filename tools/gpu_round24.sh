#!/bin/bash
# 2- and 4-GPU weak-scaling lines of cfg3 (run on a 4-GPU box)
mkdir -p gpurun_out
for n in 2 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline \
      > gpurun_out/bench_cfg3_n$n.json 2> gpurun_out/bench_cfg3_n$n.err
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/bench_cfg3_n{n}.json").read().splitlines() if l.startswith("{")][-1])
    print("cfg3 N =", n, "ms/step", round(d["ms_per_step"], 2), "value", round(d["value"] / 1e6, 1), "M  e2e", round(d["e2e"]["value"] / 1e6, 1), "M allreduce us", d["allreduce_us"], d["loss_check"]["ok"], d["clocks"])
except Exception as e:
    print("FAILED", e); print(open(f"gpurun_out/bench_cfg3_n{n}.err").read()[-1200:])
PY
done
