#!/bin/bash
# One GPU-box visit: parity tests, launch list, full ncu captures of the two MU kernels, short bench.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/prof_run.py 148 2.0 > gpurun_out/prof_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hsolve_fast -s 60 -c 1 -f -o gpurun_out/hsolve_fast python tools/prof_run.py 148 2.0 > gpurun_out/prof_h.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wsolve_fast -s 60 -c 1 -f -o gpurun_out/wsolve_fast python tools/prof_run.py 148 2.0 > gpurun_out/prof_w.log 2>&1
python bench.py --steps 2 --warmup 3 > gpurun_out/bench.log 2>&1
tail -2 gpurun_out/bench.log
