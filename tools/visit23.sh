mkdir -p gpurun_out
for cfg in "c8 c8 3" "c8 c8 4" "c8 c8 6" "c8 c4 3" "c4 c8 3"; do
set -- $cfg
SNMFNAT_HSOLVE=$1 SNMFNAT_WSOLVE=$2 timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --groups $3 > gpurun_out/bench_mix.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_mix.log") if x.startswith("{")]
d=json.loads(l[-1]); print("h=$1 w=$2 groups=$3 xRT", round(d["value"],1), "ms", round(d["ms_per_step"],1), "| serial h", round(d["roofline_all"]["hsolve"]["ms_per_step"]), "w", round(d["roofline_all"]["wsolve"]["ms_per_step"]))
PY
done
