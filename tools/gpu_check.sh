#!/bin/bash
# One GPU-box visit: parity tests, then the default bench line (headline + oracle spot check + training leg).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 1500 gpurun_out/bench.err; tail -c 600 gpurun_out/bench.log
