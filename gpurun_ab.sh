L=st-mgcn_b200/lib
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['ms_per_step'], d['value'], d['roofline_lstm']['forward']['ms'], d['roofline_lstm']['backward']['ms'])"; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
run new
cp $L/base.so $L/libstmgcn_b200.so; run base
cp $L/new.so $L/libstmgcn_b200.so; run new
cp $L/base.so $L/libstmgcn_b200.so; run base
cp $L/new.so $L/libstmgcn_b200.so
