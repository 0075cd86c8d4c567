#!/bin/bash
# round-2 GPU call H: tests, c5 with 2 / 3 resident solve CTAs, c3s, default bench
mkdir -p gpurun_out/r02h
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02h/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02h/pytest.log
tail -4 gpurun_out/r02h/pytest.log
for mb in 2 3; do
PPCA_B200_SOLVE64_MINB=$mb timeout 900 python bench.py --workload c5 --steps 5 > gpurun_out/r02h/bench_c5_minb$mb.json 2> gpurun_out/r02h/bench_c5_minb$mb.err
python - <<PY
import json
j=json.loads(open("gpurun_out/r02h/bench_c5_minb$mb.json").read().strip().splitlines()[-1])
print("c5 minb=$mb", j["ms_per_step"], j["value"], j["roofline"]["family_ms_per_step"])
PY
done
timeout 900 python bench.py --workload c3s --steps 5 --no-cpu --no-blocks > gpurun_out/r02h/bench_c3s.json 2> gpurun_out/r02h/bench_c3s.err
python - <<PY
import json
j=json.loads(open("gpurun_out/r02h/bench_c3s.json").read().strip().splitlines()[-1])
print("c3s", j["ms_per_step"], j["value"], j["roofline"]["family_ms_per_step"])
PY
