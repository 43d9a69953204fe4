#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/profile_bake_teaser.py 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_bake_launches_teaser.csv python scripts/profile_bake_teaser.py > gpurun_out/bake_ncu.log 2>&1; echo "ncu exit $?"
python - <<'PY'
import csv,collections
rows=list(csv.reader(l for l in open('gpurun_out/r01_bake_launches_teaser.csv') if l.startswith('"')))
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg={}; cnt=collections.Counter()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    k=r[ki][:48]; agg[k]=agg.get(k,0)+v; cnt[k]+=1
for k,v in sorted(agg.items(), key=lambda x:-x[1])[:10]: print(f"{k:48s} {cnt[k]:4d} {v/cnt[k]/1e3:9.1f} us each")
PY
