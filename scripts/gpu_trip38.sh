#!/bin/bash
# launch list of the final round-1 step (same command as the bench, side measurements off): 2 timed steps after 3 warm-up steps
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 950 -c 640 --csv --log-file gpurun_out/r01_launches_final2.csv python bench.py --steps 2 --warmup 3 --no-bake --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu exit $?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(l for l in open('gpurun_out/r01_launches_final2.csv') if l.startswith('"')))
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg={}; cnt=collections.Counter()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ki][:60]; agg[k]=agg.get(k,0)+v; cnt[k]+=1
tot=sum(agg.values())
for k,v in sorted(agg.items(), key=lambda x:-x[1])[:8]: print(f"{k:60s} {cnt[k]:4d} {v/cnt[k]/1e3:9.1f} us each {100*v/tot:5.1f}%")
PY
