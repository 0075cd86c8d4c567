#!/bin/bash
# round-2 GPU call C: GPU tests (single-pass mixture, sharded C ABI), default bench line with the c3/c4 blocks
mkdir -p gpurun_out/r02c
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02c/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c/pytest.log
tail -5 gpurun_out/r02c/pytest.log
timeout 900 python bench.py > gpurun_out/r02c/bench_c2.json 2> gpurun_out/r02c/bench_c2.err; echo "bench exit $?"
tail -3 gpurun_out/r02c/bench_c2.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02c/bench_c2.json").read().strip().splitlines()[-1])
print("c2", j["ms_per_step"], j["value"], j["roofline"]["family_ms_per_step"])
print("parity", {k:v for k,v in j["parity"].items() if k!="against"})
print("strong", j["strong_scaling"])
for b in ("c3_shard","c4_shard"):
    x=j[b]; print(b, x["ms_per_step"], x["value"], {k:round(v["ms_per_step"],3) for k,v in x["families"].items()}, x.get("contraction"))
print("e2e", j["e2e"]["value"], j["kernel_variants"])
PY
