run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['ms_per_step'], d['value'], d['e2e'])"; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run streams1
STMGCN_GRAPH_STREAMS=0 run streams0
run streams1
STMGCN_GRAPH_STREAMS=0 run streams0
