#!/bin/bash
# final check after the last kernel change: whole GPU suite, memcheck of the row kernels with the prefetch, default bench line
O=gpurun_out/r02final2; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
SANITIZE_ONLY=0,1,2 PPCA_B200_SOLVE_PREFETCH=1 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > $O/memcheck_prefetch.log 2>&1; echo "memcheck exit $?" | tee -a $O/memcheck_prefetch.log
SANITIZE_ONLY=0,1,2 PPCA_B200_SOLVE_PREFETCH=1 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py > $O/racecheck_prefetch.log 2>&1; echo "racecheck exit $?" | tee -a $O/racecheck_prefetch.log
tail -2 $O/memcheck_prefetch.log; tail -2 $O/racecheck_prefetch.log
python3 bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench exit $?"
python - <<PY
import json
j=json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
print("c2", round(j["ms_per_step"],3), "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "parity", j["parity"]["ok"])
for b in ("strong_scaling","c3_shard","c4_shard","c3_full"):
    x=j.get(b)
    if x: print(b, round(x["ms_per_step"],2), round(x["value"]))
PY
