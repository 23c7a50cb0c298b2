mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.log 2>&1
grep '^{' gpurun_out/bench_n2.log | tail -1 | cut -c1-330
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --steps 10 --warmup 3 > gpurun_out/bench_train_n2_weak.log 2>&1
grep '^{' gpurun_out/bench_train_n2_weak.log | tail -1 | cut -c1-330
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload train --train-scaling strong --train-frames 2500000 --steps 10 --warmup 3 > gpurun_out/bench_train_n2_strong.log 2>&1
grep '^{' gpurun_out/bench_train_n2_strong.log | tail -1 | cut -c1-330
