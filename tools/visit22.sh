set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wsolve_fast -s 60 -c 1 -f -o gpurun_out/wsolve_fast python tools/prof_run.py 148 2.0 > gpurun_out/prof_w.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hsolve_fast -s 60 -c 1 -f -o gpurun_out/hsolve_fast python tools/prof_run.py 148 2.0 > gpurun_out/prof_h.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --utts 256 > gpurun_out/b_ncu.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1
tail -c 400 gpurun_out/bench_default.log
