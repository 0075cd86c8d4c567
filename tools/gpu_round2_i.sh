#!/bin/bash
# round-2 GPU call I: tests + default bench + c3s (persistent proj kernel, 4 resident cross-moment CTAs)
mkdir -p gpurun_out/r02i
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02i/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i/pytest.log
tail -4 gpurun_out/r02i/pytest.log
timeout 900 python bench.py --no-cpu --no-blocks > gpurun_out/r02i/bench_c2.json 2> gpurun_out/r02i/bench_c2.err
python - <<PY
import json
j=json.loads(open("gpurun_out/r02i/bench_c2.json").read().strip().splitlines()[-1])
print("c2", j["ms_per_step"], j["value"], j["roofline"]["family_ms_per_step"])
PY
for mb in 2 3; do
PPCA_B200_CROSS_MINB=$mb timeout 900 python bench.py --no-cpu --no-blocks > gpurun_out/r02i/bench_c2_cross$mb.json 2> gpurun_out/r02i/bench_c2_cross$mb.err
python - <<PY
import json
j=json.loads(open("gpurun_out/r02i/bench_c2_cross$mb.json").read().strip().splitlines()[-1])
print("c2 cross minb=$mb", j["ms_per_step"], j["roofline"]["family_ms_per_step"]["cross_resid"])
PY
done
for mb in 2 4; do
PPCA_B200_SOLVE_MINB=$mb timeout 900 python bench.py --no-cpu --no-blocks > gpurun_out/r02i/bench_c2_solve$mb.json 2> gpurun_out/r02i/bench_c2_solve$mb.err
python - <<PY
import json
j=json.loads(open("gpurun_out/r02i/bench_c2_solve$mb.json").read().strip().splitlines()[-1])
print("c2 solve minb=$mb", j["ms_per_step"], j["roofline"]["family_ms_per_step"]["solve"])
PY
done
timeout 900 python bench.py --workload c3s --steps 5 --no-cpu --no-blocks > gpurun_out/r02i/bench_c3s.json 2> gpurun_out/r02i/bench_c3s.err
python - <<PY
import json
j=json.loads(open("gpurun_out/r02i/bench_c3s.json").read().strip().splitlines()[-1])
print("c3s", j["ms_per_step"], j["value"], j["roofline"]["family_ms_per_step"])
PY
