mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -m gpu -x -q 2>&1 | tail -5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/train_launches.csv python bench.py --workload train --steps 2 --warmup 1 > gpurun_out/b_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/train_launches.csv', errors='ignore')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    try: v=float(r[mv].replace(',',''))
    except: continue
    n=r[kn][:60]; agg[n][0]+=1; agg[n][1]+=v
for n,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:10]: print(f"{t/1e6:10.3f} ms total  {c:5d} launches  {t/c/1e3:10.1f} us avg  {n}")
PY
timeout 900 python bench.py --workload train --train-frames 10000000 --steps 5 --warmup 3 > gpurun_out/bench_train_10M.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench_train_10M.log") if x.startswith("{")]
if l:
    d=json.loads(l[-1]); print("10M: iters/s", round(d["value"],2), "ms", round(d["ms_per_step"],2), "TF", round(d["roofline"]["achieved"],1), "frac", round(d["roofline"]["frac"],3), "e2e", d["e2e"]["value"])
else:
    print(open("gpurun_out/bench_train_10M.log").read()[-2000:])
PY
