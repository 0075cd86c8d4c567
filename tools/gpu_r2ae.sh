#!/bin/bash
# round-2 GPU call AE: same-box A/B of two builds (PPCA_B200_LIB): tiled solve with / without the cp.async prefetch
O=gpurun_out/r02ae; mkdir -p $O
for rep in 1 2; do
for lib in prev new; do
  if [ $lib = prev ]; then export PPCA_B200_LIB=$PWD/ppca_rs_b200/libppca_b200_prev.so; else unset PPCA_B200_LIB; fi
  timeout 600 python bench.py --workload c3s --steps 5 --no-cpu --no-blocks > $O/c3s_$lib.json 2> $O/c3s_$lib.err
  timeout 600 python bench.py --workload c5 --steps 3 --no-cpu --no-blocks > $O/c5_$lib.json 2> $O/c5_$lib.err
  python - <<PY
import json
for wl in ("c3s","c5"):
    j=json.loads(open("$O/%s_$lib.json"%wl).read().strip().splitlines()[-1])
    print("$lib rep$rep", wl, "ms/step", round(j["ms_per_step"],2), "solve", round(j["roofline"]["family_ms_per_step"]["solve"],2))
PY
done
done
