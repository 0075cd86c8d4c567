#!/bin/bash
# round-2 GPU call E: GPU tests, default bench, c4 probe with the per-component ladder
mkdir -p gpurun_out/r02e
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02e/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e/pytest.log
tail -5 gpurun_out/r02e/pytest.log
timeout 900 python bench.py > gpurun_out/r02e/bench_c2.json 2> gpurun_out/r02e/bench_c2.err; echo "bench exit $?"
tail -3 gpurun_out/r02e/bench_c2.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02e/bench_c2.json").read().strip().splitlines()[-1])
print("c2", j["ms_per_step"], j["value"], j["roofline"]["family_ms_per_step"])
print("parity", {k:v for k,v in j["parity"].items() if k not in ("against","vs_reference_order")})
for b in ("c3_shard","c4_shard"):
    x=j[b]; print(b, x["ms_per_step"], x["value"], {k:round(v["ms_per_step"],3) for k,v in x["families"].items()})
print("e2e", j["e2e"]["value"], j["kernel_variants"])
PY
PPCA_B200_DEBUG=1 timeout 600 python bench.py --workload c4 --rows 131072 --steps 3 --no-cpu --no-blocks > gpurun_out/r02e/bench_c4_probe.json 2> gpurun_out/r02e/bench_c4_probe.err
grep "ppca_b200" gpurun_out/r02e/bench_c4_probe.err | sort | uniq -c | head
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02e/bench_c4_probe.json").read().strip().splitlines()[-1])
print("c4 probe", j["ms_per_step"], j["roofline"]["family_ms_per_step"], j["kernel_variants"], j["gpu_launches"])
PY
