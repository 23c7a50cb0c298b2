set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_parity.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_parity.log
tail -15 gpurun_out/pytest_parity.log
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_reg.log 2>&1
tail -2 gpurun_out/bench_reg.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hsolve_reg -s 60 -c 1 -f -o gpurun_out/hsolve_reg python tools/prof_run.py 148 2.0 > gpurun_out/prof_hr.log 2>&1
