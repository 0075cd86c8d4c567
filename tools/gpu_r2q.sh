#!/bin/bash
# round-2 GPU call Q: launch list of one c4 (PPCAMix M=32) step: which kernels, how many, how long (ncu is serialising: shares only)
O=gpurun_out/r02q; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c4.csv \
  python bench.py --workload c4 --rows 65536 --steps 1 --warmup 1 --no-cpu --no-blocks > $O/ncu_c4.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$O/launches_c4.csv")) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
cnt=collections.Counter(); tim=collections.Counter()
for r in rows[1:]:
    name=r[ix["Kernel Name"]].split("(")[0][:60]
    cnt[name]+=1
    try: tim[name]+=float(r[ix["Metric Value"]].replace(",",""))
    except: pass
tot=sum(tim.values())
print("launches", sum(cnt.values()), "total us", tot/1e3)
for n,c in cnt.most_common(40): print(f"{c:6d} {tim[n]/1e3:10.1f} us {tim[n]/c/1e3:8.2f} us/launch  {n}")
PY
