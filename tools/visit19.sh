mkdir -p gpurun_out
SNMFNAT_TRAIN_DEBUG=1 timeout 300 python tools/train_prof_run.py 2>&1 | grep -i "probe\|error\|Traceback\|timeout" | head
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -25
for mc in 1 0; do
SNMFNAT_TRAIN_MC=$mc timeout 600 python bench.py --workload train --steps 5 --warmup 3 > gpurun_out/bench_train_mc$mc.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_train_mc$mc.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("mc=$mc iters/s", round(d["value"],2), "ms", round(d["ms_per_step"],2), "TF", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3))
else:
    print(open("gpurun_out/bench_train_mc$mc.log").read()[-2000:])
PY
done
