L=st-mgcn_b200/lib
run() { python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['ms_per_step'], d['value'])"; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
run new
cp $L/base.so $L/libstmgcn_b200.so; run base
cp $L/new.so $L/libstmgcn_b200.so; run new
cp $L/base.so $L/libstmgcn_b200.so; run base
cp $L/new.so $L/libstmgcn_b200.so
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "tensor_core or golden" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
