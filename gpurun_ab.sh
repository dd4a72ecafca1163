mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r10_n8.json 2> gpurun_out/bench_r10_n8.err
tail -c 300 gpurun_out/bench_r10_n8.json
