#!/bin/bash
# One GPU-box visit: parity tests, then a short serialised per-class profile and a short bench of the enhancement path.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python tools/prof_run.py 1024 3.0 > gpurun_out/prof_ms.log 2>&1; tail -3 gpurun_out/prof_ms.log
SNMFNAT_HSOLVE=single timeout 300 python tools/prof_run.py 1024 3.0 > gpurun_out/prof_single.log 2>&1; tail -3 gpurun_out/prof_single.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log
