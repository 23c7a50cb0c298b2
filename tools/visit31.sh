mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1
grep '^{' gpurun_out/bench_default.log | tail -1 | cut -c1-260
