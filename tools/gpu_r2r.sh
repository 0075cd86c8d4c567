#!/bin/bash
# round-2 GPU call R: CUDA-graph replay of the mixture chunk loop: parity, then c4 shard timing graphs on / off
O=gpurun_out/r02r; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "mixture or golden_mixture or sharded_entry_points_world or dmma_mixture" > $O/pytest_mix.log 2>&1; echo "pytest exit $?" >> $O/pytest_mix.log
tail -15 $O/pytest_mix.log
for g in 1 0; do
  PPCA_B200_GRAPHS=$g timeout 600 python bench.py --workload c4 --rows 131072 --steps 3 --no-cpu --no-blocks > $O/bench_c4_g$g.json 2> $O/bench_c4_g$g.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/bench_c4_g$g.json").read().strip().splitlines()[-1])
    print("graphs=$g c4 ms/step", round(j["ms_per_step"],2), "launches", j["gpu_launches"], "variants", j.get("kernel_variants"), "fam", {k:round(v,1) for k,v in j["roofline"]["family_ms_per_step"].items()})
except Exception as e:
    print("graphs=$g failed", e); print(open("$O/bench_c4_g$g.err").read()[-1500:])
PY
done
nproc; lscpu | grep -E "Model name|MHz" | head -3
