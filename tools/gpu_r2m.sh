#!/bin/bash
# round-2 GPU call M: register-tiled solve kernel — parity tests first, then old-vs-new timings at the c3 / c4 / c5 shapes
O=gpurun_out/r02m; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "state_size_sweep or tiled_solve or test_infer or test_llks or iterate_trajectory or two_tile or mixture" > $O/pytest_solve.log 2>&1; echo "pytest exit $?" >> $O/pytest_solve.log
tail -15 $O/pytest_solve.log
for mode in tile rows; do
  export PPCA_B200_SOLVE=$mode
  timeout 600 python bench.py --workload c3s --steps 5 --no-cpu --no-blocks > $O/bench_c3s_$mode.json 2> $O/bench_c3s_$mode.err
  timeout 600 python bench.py --workload c5 --steps 3 --no-cpu --no-blocks > $O/bench_c5_$mode.json 2> $O/bench_c5_$mode.err
  timeout 600 python bench.py --workload c4 --rows 131072 --steps 3 --no-cpu --no-blocks > $O/bench_c4_$mode.json 2> $O/bench_c4_$mode.err
  python - <<PY
import json
for wl in ("c3s","c5","c4"):
    try:
        j=json.loads(open("$O/bench_%s_$mode.json"%wl).read().strip().splitlines()[-1])
        f=j["roofline"].get("family_ms_per_step") or {}
        print("$mode", wl, "ms/step", round(j["ms_per_step"],2), "solve", round(f.get("solve",0),2), "variants", j.get("kernel_variants"))
    except Exception as e:
        print("$mode", wl, "failed", e)
PY
done
