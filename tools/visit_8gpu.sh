mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n8.log 2>&1
grep '^{' gpurun_out/bench_n8.log | tail -1 | cut -c1-330
tail -3 gpurun_out/bench_n8.log | cut -c1-300
