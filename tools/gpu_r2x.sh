#!/bin/bash
# round-2 GPU call X: compute-sanitizer over the final kernels (the panel-blocked solve changed after the last run), the
# launch list of one c5 inference step, smoke()
O=gpurun_out/r02x; mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > $O/memcheck.log 2>&1; echo "memcheck exit $?" | tee -a $O/memcheck.log
tail -4 $O/memcheck.log
SANITIZE_ONLY=2,3,x timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py > $O/racecheck.log 2>&1; echo "racecheck exit $?" | tee -a $O/racecheck.log
tail -4 $O/racecheck.log
PPCA_B200_SOLVE16=tile PPCA_B200_SOLVE=tile SANITIZE_ONLY=0,2 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py > $O/racecheck_tile_small_k.log 2>&1; echo "racecheck (tile, k <= 32) exit $?" | tee -a $O/racecheck_tile_small_k.log
tail -3 $O/racecheck_tile_small_k.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_c5.csv \
  python bench.py --workload c5 --rows 500000 --steps 1 --warmup 1 --no-cpu --no-blocks > $O/launches_c5.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$O/launches_c5.csv")) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
cnt=collections.Counter(); tim=collections.Counter()
for r in rows[1:]:
    name=r[ix["Kernel Name"]].split("(")[0][:60]
    cnt[name]+=1
    try: tim[name]+=float(r[ix["Metric Value"]].replace(",",""))
    except: pass
for n,c in sorted(cnt.items(), key=lambda kv:-tim[kv[0]])[:14]: print(f"{c:5d} {tim[n]/1e3:10.1f} us {tim[n]/c/1e3:9.1f} us/launch  {n}")
PY
python -c "import __graft_entry__ as g; g.smoke()"
