#!/bin/bash
# One GPU-box visit: parity tests, then the default bench line (headline + oracle spot check + training leg) and the CPU arm.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 1500 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.err; tail -c 300 gpurun_out/bench_default.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -c 300 gpurun_out/bench_reference.log
