#!/bin/bash
# round-2 GPU call S: whole GPU suite, compute-sanitizer memcheck + racecheck over every kernel variant (incl. the later
# round-2 kernels), c5 / c3s benches with the shipped solve dispatch
O=gpurun_out/r02s; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > $O/memcheck.log 2>&1; echo "memcheck exit $?" | tee -a $O/memcheck.log
tail -6 $O/memcheck.log
SANITIZE_ONLY=2,3,x timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py > $O/racecheck.log 2>&1; echo "racecheck exit $?" | tee -a $O/racecheck.log
tail -5 $O/racecheck.log
timeout 600 python bench.py --workload c5 --steps 3 --no-cpu --no-blocks > $O/bench_c5.json 2> $O/bench_c5.err
timeout 600 python bench.py --workload c3s --steps 5 --no-cpu --no-blocks > $O/bench_c3s.json 2> $O/bench_c3s.err
python - <<PY
import json
for wl in ("c5","c3s"):
    j=json.loads(open("$O/bench_%s.json"%wl).read().strip().splitlines()[-1])
    print(wl, "ms/step", round(j["ms_per_step"],2), "value", round(j["value"]), "e2e", round(j["e2e"]["value"]) if j.get("e2e") else None, j["roofline"].get("family_ms_per_step"))
PY
