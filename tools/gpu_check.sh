#!/bin/bash
# One GPU-box visit: parity tests, launch lists, full ncu captures of the MU kernels (online + training), short benches.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 2 --warmup 3 > gpurun_out/bench.log 2>&1
tail -2 gpurun_out/bench.log
python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/bench_train.log 2>&1
tail -2 gpurun_out/bench_train.log
python bench.py --workload train --train-frames 10000000 --steps 5 --warmup 3 > gpurun_out/bench_train_10M.log 2>&1
tail -2 gpurun_out/bench_train_10M.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/train_launches.csv python tools/train_prof_run.py > gpurun_out/prof_train_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hphase -s 1 -c 1 -f -o gpurun_out/train_hphase python tools/train_prof_run.py > gpurun_out/prof_th.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wphase -s 1 -c 1 -f -o gpurun_out/train_wphase python tools/train_prof_run.py > gpurun_out/prof_tw.log 2>&1
