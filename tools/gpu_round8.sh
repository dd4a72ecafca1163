#!/bin/bash
# 8-GPU bench lines: cfg3 weak scaling (fp32-grade), cfg4 (configs[3]: batch 512 bf16), cfg5 (configs[4]: batch 256 bf16)
mkdir -p gpurun_out
run() { # name, args...
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 "$@" \
      > gpurun_out/$name.json 2> gpurun_out/$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads([l for l in open(f"gpurun_out/{name}.json").read().splitlines() if l.startswith("{")][-1])   # (NCCL may print a banner first)
    print(name, "ms/step", round(d["ms_per_step"], 2), "value", round(d["value"] / 1e6, 1), "M  e2e", round(d["e2e"]["value"] / 1e6, 1),
          "M  per-rank ms", [round(v, 2) for v in d["per_rank_ms_per_step"]], "allreduce us", d["allreduce_us"], d["loss_check"]["ok"], d["clocks"])
except Exception as e:
    print(name, "FAILED", e)
    print(open(f"gpurun_out/{name}.err").read()[-1500:])
PY
}
run bench_cfg3_n8 --steps 10 --warmup 3
run bench_cfg4_bf16_n8 --workload cfg4 --steps 10 --warmup 3
run bench_cfg5_bf16_n8 --workload cfg5 --steps 4 --warmup 3
