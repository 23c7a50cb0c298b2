set -x
mkdir -p gpurun_out
timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_parity.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_parity.log
tail -15 gpurun_out/pytest_parity.log
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_cur.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_cur.log") if x.startswith("{")]
d=json.loads(l[-1]); print("xRT", round(d["value"],1), "ms", round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],1))
for k,v in d["roofline_all"].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ("ms_per_step","frac","achieved")})
PY
