mkdir -p gpurun_out
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/bench_r8_n$n.json 2> gpurun_out/bench_r8_n$n.err
tail -c 600 gpurun_out/bench_r8_n$n.json | cut -c1-400
done
