#!/bin/bash
# Two-GPU visit: the whole GPU suite (the NCCL test needs 2 devices), then the 2-rank bench line (weak + strong + training).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
grep -E "passed|failed|error|60 iterations|rc=" gpurun_out/pytest_gpu2.log | tail -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; tail -c 800 gpurun_out/bench_n2.err; tail -c 300 gpurun_out/bench_n2.log
