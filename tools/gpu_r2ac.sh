#!/bin/bash
# round-2 GPU call AC: cp.async prefetch of the next group's packed rows in the lane-owns-a-row solve kernels (k <= 32)
O=gpurun_out/r02ac; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "state_size_sweep or test_infer or test_llks or iterate_trajectory or test_mixture or golden or precision_guard or streaming" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -4 $O/pytest.log
for pf in 1 0; do
  PPCA_B200_SOLVE_PREFETCH=$pf timeout 600 python bench.py --no-cpu --no-blocks > $O/bench_c2_pf$pf.json 2> $O/bench_c2_pf$pf.err
  PPCA_B200_SOLVE_PREFETCH=$pf timeout 600 python bench.py --workload c4 --rows 131072 --steps 3 --no-cpu --no-blocks > $O/bench_c4_pf$pf.json 2> $O/bench_c4_pf$pf.err
  python - <<PY
import json
for wl in ("c2","c4"):
    try:
        j=json.loads(open("$O/bench_%s_pf$pf.json"%wl).read().strip().splitlines()[-1])
        print("prefetch=$pf", wl, "ms/step", round(j["ms_per_step"],3), "solve", round(j["roofline"]["family_ms_per_step"]["solve"],3))
    except Exception as e:
        print("prefetch=$pf", wl, "failed", e); print(open("$O/bench_%s_pf$pf.err"%wl).read()[-600:])
PY
done
