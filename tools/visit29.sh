mkdir -p gpurun_out
timeout 600 python bench.py --workload train --train-k 100 --steps 10 --warmup 3 > gpurun_out/bench_train_k100.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_train_k100.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("K=100: iters/s", round(d["value"],2), "ms", round(d["ms_per_step"],2), "TF", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3))
else:
    print(open("gpurun_out/bench_train_k100.log").read()[-1500:])
PY
SNMFNAT_TRAIN_V1=1 timeout 600 python bench.py --workload train --train-k 100 --steps 10 --warmup 3 > gpurun_out/bench_train_k100_v1.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_train_k100_v1.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("K=100 gen1: iters/s", round(d["value"],2), "ms", round(d["ms_per_step"],2), "TF", round(d["roofline"]["achieved"],1))
PY
