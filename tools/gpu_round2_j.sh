#!/bin/bash
# round-2 GPU call J: tests + default bench with the compact host format in the e2e leg (+ the plain format for comparison)
mkdir -p gpurun_out/r02j
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02j/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02j/pytest.log
tail -4 gpurun_out/r02j/pytest.log
timeout 900 python bench.py > gpurun_out/r02j/bench_c2.json 2> gpurun_out/r02j/bench_c2.err; echo "bench exit $?"
tail -2 gpurun_out/r02j/bench_c2.err
python - <<PY
import json
j=json.loads(open("gpurun_out/r02j/bench_c2.json").read().strip().splitlines()[-1])
print("c2", j["ms_per_step"], j["value"], "e2e", j["e2e"]["value"], j["e2e"]["h2d_gb_per_s"], j["e2e"]["host_bytes_per_sample"], j["e2e"]["call"])
print("cpu", j["cpu_baseline"]["value"], j["cpu_baseline"]["cores"], "parity ok", j["parity"]["ok"])
PY
timeout 900 python bench.py --no-cpu --no-blocks --e2e-plain > gpurun_out/r02j/bench_c2_plain.json 2> gpurun_out/r02j/bench_c2_plain.err
python - <<PY
import json
j=json.loads(open("gpurun_out/r02j/bench_c2_plain.json").read().strip().splitlines()[-1])
print("c2 plain e2e", j["e2e"]["value"], j["e2e"]["h2d_gb_per_s"])
PY
timeout 900 python bench.py --workload c3s --steps 5 --no-cpu --no-blocks > gpurun_out/r02j/bench_c3s.json 2> gpurun_out/r02j/bench_c3s.err
python - <<PY
import json
j=json.loads(open("gpurun_out/r02j/bench_c3s.json").read().strip().splitlines()[-1])
print("c3s", j["ms_per_step"], j["value"], "e2e", j["e2e"]["value"], j["e2e"]["h2d_gb_per_s"], j["e2e"]["host_bytes_per_sample"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02j/bench_c2_reference.json 2> gpurun_out/r02j/bench_c2_reference.err
tail -c 600 gpurun_out/r02j/bench_c2_reference.json
