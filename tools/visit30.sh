mkdir -p gpurun_out
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4
for v in 1 0; do
SNMFNAT_HFUSE=$v timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_fuse.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_fuse.log") if x.startswith("{")]
d=json.loads(l[-1]); print("fuse=$v xRT", round(d["value"],1), "ms", round(d["ms_per_step"],1), "hsolve", round(d["roofline_all"]["hsolve"]["ms_per_step"]), "checksum", d["output_checksum"])
PY
done
