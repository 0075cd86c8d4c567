#!/bin/bash
# final check of the round: what the driver runs — both bench arms with the driver's flags (tests + smoke ran in the call before)
O=gpurun_out/r02final; mkdir -p $O
T0=$(date +%s); python3 bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err; echo "reference arm exit $? in $(( $(date +%s) - T0 )) s"
T0=$(date +%s); python3 bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err; echo "bench exit $? in $(( $(date +%s) - T0 )) s"
python - <<PY
import json
j=json.loads(open("$O/bench.json").read().strip().splitlines()[-1])
r=json.loads(open("$O/bench_reference.json").read().strip().splitlines()[-1])
print("c2", round(j["ms_per_step"],3), "value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "ref", round(r["value"]), "e2e ratio", round(j["e2e"]["value"]/r["value"],1), "parity", j["parity"]["ok"], "launches", j["gpu_launches"], "clocks", j["clocks"])
for b in ("strong_scaling","c3_shard","c4_shard","c3_full"):
    x=j.get(b)
    if x: print(b, round(x["ms_per_step"],2), round(x["value"]))
PY
