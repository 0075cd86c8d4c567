#!/bin/bash
# round-2 GPU call P: fast generator (Xi -> DMMA row GEMM -> noise + mask), GeneratedDataset parity, default bench with c3_full
O=gpurun_out/r02p; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "generated_dataset or model_sample or reference_toy or reference_big or examples or full_size or samplers or dataset_roundtrip" > $O/pytest_gen.log 2>&1; echo "pytest exit $?" >> $O/pytest_gen.log
tail -12 $O/pytest_gen.log
timeout 1500 python bench.py > $O/bench_c2.json 2> $O/bench_c2.err; echo "bench exit $?"
tail -3 $O/bench_c2.err
python - <<PY
import json
j=json.loads(open("$O/bench_c2.json").read().strip().splitlines()[-1])
print("c2", j["ms_per_step"], j["value"], "e2e", j["e2e"]["value"], "parity", j["parity"])
for b in ("c3_shard","c4_shard","c3_full"):
    x=j.get(b)
    if x: print(b, x["rows_per_gpu"], round(x["ms_per_step"],2), round(x["value"]), {k:round(v,2) for k,v in (x.get("family_ms_per_step") or {}).items()}, x.get("generator_and_ingest_ms_per_step"))
PY
