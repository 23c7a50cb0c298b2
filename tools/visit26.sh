mkdir -p gpurun_out
for g in 2 3 4 5; do
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --groups $g > gpurun_out/bench_g.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_g.log") if x.startswith("{")]
d=json.loads(l[-1]); print("groups=$g xRT", round(d["value"],1), "ms", round(d["ms_per_step"],1))
PY
done
