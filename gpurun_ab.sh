run() { python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['ms_per_step'], d['value'])"; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run tma1
STMGCN_BWD_TMA=0 run tma0
run tma1
STMGCN_BWD_TMA=0 run tma0
