#!/bin/bash
# round-2 GPU call AB (8 GPUs): the driver's scaling command at N = 8
O=gpurun_out/r02ab; mkdir -p $O
nvidia-smi -L | wc -l
T0=$(date +%s)
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus ${NG:-8} --steps 20 --warmup 5 > $O/bench_c2_${NG:-8}gpu.json 2> $O/bench_c2_${NG:-8}gpu.err; echo "bench exit $? in $(( $(date +%s) - T0 )) s"
tail -3 $O/bench_c2_${NG:-8}gpu.err
python - <<PY
import json
j=json.loads(open("$O/bench_c2_${NG:-8}gpu.json").read().strip().splitlines()[-1])
print("c2 x${NG:-8}", j["ms_per_step"], j["value"], "e2e", j["e2e"]["value"], j["e2e"].get("h2d_gb_per_s"), "strong", j["strong_scaling"]["value"] if j.get("strong_scaling") else None)
for b in ("c3_shard","c4_shard","c3_full"):
    x=j.get(b)
    if x: print(b, x["rows_per_gpu"], round(x["ms_per_step"],2), round(x["value"]), "comm", x.get("comm_ms_per_step"), x.get("kernel_variants"))
PY
