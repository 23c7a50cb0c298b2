mkdir -p gpurun_out
SNMFNAT_HPASSB=1 timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
for v in 0 1; do
SNMFNAT_HPASSB=$v timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_pb.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_pb.log") if x.startswith("{")]
d=json.loads(l[-1]); print("passB=$v xRT", round(d["value"],1), "ms", round(d["ms_per_step"],1), "hsolve", round(d["roofline_all"]["hsolve"]["ms_per_step"]))
PY
done
