set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_parity.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_parity.log
tail -5 gpurun_out/pytest_parity.log
for g in 1 2 3 4 6; do
  timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --groups $g > gpurun_out/bench_g$g.log 2>&1
  python - <<PY
import json
l=[x for x in open("gpurun_out/bench_g$g.log") if x.startswith("{")]
d=json.loads(l[-1]); print("groups", $g, "xRT", round(d["value"],1), "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1))
PY
done
