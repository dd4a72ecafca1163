#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_graphs.py -m gpu -q -x 2>&1 | tail -3 | cut -c1-300
for wl in cfg1 cfg2; do for g in "" "--cuda-graph"; do
  python bench.py --workload $wl --arith fp32 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e $g 2>gpurun_out/graph_$wl.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl', '$g', 'ms/step', round(d['ms_per_step'],3), 'M r-t/s', round(d['value']/1e6,2), d['loss_check']['ok'], d['gpu_launches'])" || tail -5 gpurun_out/graph_$wl.err
done; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --cuda-graph 2>gpurun_out/graph_cfg3.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 --cuda-graph ms/step', round(d['ms_per_step'],3), 'M r-t/s', round(d['value']/1e6,2), d['loss_check']['ok'])" || tail -5 gpurun_out/graph_cfg3.err
