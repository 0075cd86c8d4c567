#!/bin/bash
# round-2 GPU call B: GPU test suite, memcheck over every kernel variant, racecheck on the remaining variants
mkdir -p gpurun_out/r02b
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02b/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b/pytest.log
tail -5 gpurun_out/r02b/pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/r02b/memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02b/memcheck.log
tail -4 gpurun_out/r02b/memcheck.log
SANITIZE_ONLY=1,3,6 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/r02b/racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r02b/racecheck.log
tail -4 gpurun_out/r02b/racecheck.log
timeout 600 python bench.py --no-cpu > gpurun_out/r02b/bench_c2.json 2> gpurun_out/r02b/bench_c2.err; echo "bench exit $?"
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02b/bench_c2.json").read().strip().splitlines()[-1])
print(j["ms_per_step"], j["roofline"]["family_ms_per_step"])
PY
