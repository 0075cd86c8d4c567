#!/bin/bash
# round-2 GPU call U: the evidence set for the final state — bench lines (default with its blocks, c3s, c4, c5, reference arm),
# launch lists and ncu --set full captures of one EM step (c2, c3 shard)
O=gpurun_out/r02u; mkdir -p $O
timeout 1500 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench c2 exit $?"
timeout 900 python bench.py --workload c3s --steps 5 --no-cpu --no-blocks > $O/bench_c3s.json 2> $O/bench_c3s.err; echo "c3s exit $?"
timeout 900 python bench.py --workload c5 --steps 3 --no-cpu --no-blocks > $O/bench_c5.json 2> $O/bench_c5.err; echo "c5 exit $?"
timeout 900 python bench.py --workload c4 --steps 3 --no-cpu --no-blocks > $O/bench_c4.json 2> $O/bench_c4.err; echo "c4 exit $?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_c2_reference_arm.json 2> $O/bench_c2_reference_arm.err; echo "ref exit $?"
python - <<PY
import json
for wl in ("c2","c3s","c5","c4","c2_reference_arm"):
    try:
        j=json.loads(open("$O/bench_%s.json"%wl).read().strip().splitlines()[-1])
        print(wl, "ms/step", round(j["ms_per_step"],3), "value", round(j["value"]), "e2e", round(j["e2e"]["value"]) if j.get("e2e") else None)
    except Exception as e:
        print(wl, "failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches_c2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-blocks > $O/launches_c2.log 2>&1; echo "launch list c2 exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/launches_c3s.csv \
  python bench.py --workload c3s --steps 2 --warmup 3 --no-cpu --no-blocks > $O/launches_c3s.log 2>&1; echo "launch list c3s exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on --launch-skip 80 -c 24 -o $O/c2_step \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-blocks > $O/ncu_c2.log 2>&1; echo "ncu c2 exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on --launch-skip 130 -c 24 -o $O/c3s_step \
  python bench.py --workload c3s --steps 2 --warmup 3 --no-cpu --no-blocks > $O/ncu_c3s.log 2>&1; echo "ncu c3s exit $?"
for f in c2_step c3s_step; do
  ncu -i $O/$f.ncu-rep --page raw --csv > $O/${f}_raw.csv 2>/dev/null
  ncu -i $O/$f.ncu-rep --page details > $O/${f}_details.txt 2>/dev/null
  ls -la $O/$f.ncu-rep
  rm -f $O/$f.ncu-rep
done
