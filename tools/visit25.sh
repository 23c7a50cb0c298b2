mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:hsolve|wsolve|gain_kernel|frame_pcm|stft_post|istft_pre|ola_int16|fft" -s 6000 -c 1800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches.csv', errors='ignore')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    try: v=float(r[mv].replace(',',''))
    except: continue
    n=r[kn][:40]; agg[n][0]+=1; agg[n][1]+=v
tot=sum(t for c,t in agg.values())
for n,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:8]: print(f"{t/1e6:10.3f} ms {100*t/tot:5.1f}%  {c:5d} launches  {n}")
PY
