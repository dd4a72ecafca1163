#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "spmm or bf16" 2>&1 | tail -3 | cut -c1-300
python tools/spmm_probe.py 2>&1 | tail -2
for a in bf16 fp32; do python bench.py --workload cfg2 --arith $a --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2', '$a', 'ms/step', round(d['ms_per_step'],3), 'M r-t/s', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), d['loss_check'])"; done
python bench.py --workload cfg5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg5 per-GPU share', d['dtype'], 'ms/step', round(d['ms_per_step'],2), 'M r-t/s', round(d['value']/1e6,2), d['loss_check'], d['config'])"
