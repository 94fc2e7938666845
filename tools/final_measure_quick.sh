#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|^FAILED" | head
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q.json 2>/dev/null; tail -1 gpurun_out/bench_q.json | cut -c1-170
ncu --metrics gpu__time_duration.sum --clock-control none -s 256 -c 256 --csv --log-file gpurun_out/launches_q.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workspace-gb 48 > /dev/null 2>&1
python - <<'PY'
import csv,collections,re
rows=list(csv.reader(open('gpurun_out/launches_q.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit')
d=collections.defaultdict(lambda:[0,0.0])
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    name=re.sub(r'\(.*','',r[kn])[:40]; v=float(r[mv].replace(',','')); u=r[mu]
    if u=='ns': v/=1e3
    elif u=='ms': v*=1e3
    d[name][0]+=1; d[name][1]+=v
for k,v in sorted(d.items(), key=lambda kv:-kv[1][1])[:12]: print('%-42s %4d %9.1f us'%(k,v[0],v[1]))
PY
