#!/bin/bash
# round-2 GPU call Z (2 GPUs): the ranks of the bench hold row ranges of ONE synthetic dataset (same truth): default line
O=gpurun_out/r02z; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "generated_dataset or two_gpus" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 > $O/bench_c2_2gpu.json 2> $O/bench_c2_2gpu.err; echo "bench exit $?"
python - <<PY
import json
j=json.loads(open("$O/bench_c2_2gpu.json").read().strip().splitlines()[-1])
print("c2 x2", j["ms_per_step"], j["value"], "e2e", j["e2e"]["value"], "strong", j["strong_scaling"]["value"] if j.get("strong_scaling") else None, "parity", j["parity"]["ok"] if j.get("parity") else None)
for b in ("c3_shard","c4_shard","c3_full"):
    x=j.get(b)
    if x: print(b, x["rows_per_gpu"], round(x["ms_per_step"],2), round(x["value"]), "comm", x.get("comm_ms_per_step"), x.get("kernel_variants"))
PY
