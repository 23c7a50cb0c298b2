mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wsolve_fast -s 60 -c 1 -f -o gpurun_out/wsolve_fast python tools/prof_run.py 148 2.0 > gpurun_out/prof_w.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1
grep '^{' gpurun_out/bench_default.log | tail -1 | cut -c1-300
