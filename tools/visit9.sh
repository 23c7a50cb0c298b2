set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q > gpurun_out/pytest_train.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_train.log
tail -5 gpurun_out/pytest_train.log
timeout 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/bench_train.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_train.log") if x.startswith("{")]
d=json.loads(l[-1]); print("iters/s", round(d["value"],2), "ms", round(d["ms_per_step"],2), "TF", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3))
PY
