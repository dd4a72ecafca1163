mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r8_n1.json 2> gpurun_out/bench_r8_n1.err; tail -c 3000 gpurun_out/bench_r8_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r8_ref.json 2> gpurun_out/bench_r8_ref.err; tail -c 1200 gpurun_out/bench_r8_ref.json
