cd $GRAFT_REPO_ROOT

timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_c5_launches.csv python bench.py --config c5 --steps 2 --warmup 2 --no-cpu --docs 20000 --init random > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_c5_launches.csv')) if len(r)>10]
h=rows[0]; ci={n:i for i,n in enumerate(h)}
from collections import defaultdict
d=defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    n=r[ci['Kernel Name']][:70]; v=float(r[ci['Metric Value']].replace(',',''))
    u=r[ci['Metric Unit']]
    v=v/1e6 if u in ('ns','nsecond') else (v/1e3 if u in ('us','usecond') else v)
    d[n][0]+=1; d[n][1]+=v
for n,(c,t) in sorted(d.items(), key=lambda kv:-kv[1][1])[:14]: print(f"{t:9.3f} ms {c:4d}x  {n}")
PY
