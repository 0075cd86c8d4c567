#!/bin/bash
# final check after the last solve change (shared column maxima, KD = 48 prefetch): GPU suite, sanitizer over the wide states, c5 / c3s lines
O=gpurun_out/r02final3; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
SANITIZE_ONLY=3,x timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > $O/memcheck_wide.log 2>&1; echo "memcheck exit $?" | tee -a $O/memcheck_wide.log
SANITIZE_ONLY=3,x timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py > $O/racecheck_wide.log 2>&1; echo "racecheck exit $?" | tee -a $O/racecheck_wide.log
tail -2 $O/memcheck_wide.log; tail -2 $O/racecheck_wide.log
timeout 600 python bench.py --workload c5 --steps 3 --no-cpu --no-blocks > $O/bench_c5.json 2> $O/bench_c5.err
timeout 600 python bench.py --workload c3s --steps 5 --no-cpu --no-blocks > $O/bench_c3s.json 2> $O/bench_c3s.err
python - <<PY
import json
for wl in ("c3s","c5"):
    j=json.loads(open("$O/bench_%s.json"%wl).read().strip().splitlines()[-1])
    print(wl, "ms/step", round(j["ms_per_step"],2), "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "solve", round(j["roofline"]["family_ms_per_step"]["solve"],2))
PY
