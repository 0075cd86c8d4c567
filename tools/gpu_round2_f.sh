#!/bin/bash
# round-2 GPU call F: GPU tests, default bench, c5 inference bench (one E-step pass, output reused in place)
mkdir -p gpurun_out/r02f
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02f/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f/pytest.log
tail -5 gpurun_out/r02f/pytest.log
timeout 900 python bench.py > gpurun_out/r02f/bench_c2.json 2> gpurun_out/r02f/bench_c2.err; echo "bench exit $?"
tail -3 gpurun_out/r02f/bench_c2.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02f/bench_c2.json").read().strip().splitlines()[-1])
print("c2", j["ms_per_step"], j["value"], j["roofline"]["family_ms_per_step"])
print("parity", {k:v for k,v in j["parity"].items() if k not in ("against","vs_reference_order")})
for b in ("c3_shard","c4_shard"):
    x=j[b]; print(b, x["ms_per_step"], x["value"], {k:round(v["ms_per_step"],3) for k,v in x["families"].items()})
print("e2e", j["e2e"]["value"], j["kernel_variants"])
PY
timeout 900 python bench.py --workload c5 --steps 5 > gpurun_out/r02f/bench_c5.json 2> gpurun_out/r02f/bench_c5.err; echo "bench c5 exit $?"
tail -3 gpurun_out/r02f/bench_c5.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02f/bench_c5.json").read().strip().splitlines()[-1])
print("c5", j["ms_per_step"], j["value"], j["roofline"]["frac"], j["roofline"]["family_ms_per_step"], "e2e", j["e2e"]["value"])
PY
