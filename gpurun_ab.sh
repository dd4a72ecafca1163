mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r10_n1.json 2> gpurun_out/bench_r10_n1.err; tail -c 300 gpurun_out/bench_r10_n1.json
