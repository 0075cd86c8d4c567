#!/bin/bash
# round-2 GPU call Y: k = 64 with 2 x 16 register tiles (the panel's publishers are a whole warp, two rows per lane)
O=gpurun_out/r02y; mkdir -p $O
PPCA_B200_SOLVE64=2x16 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "state_size_sweep or tiled_solve or two_tile or test_infer or test_llks or iterate_trajectory or smooth_extrapolate" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -5 $O/pytest.log
for mode in 2x16 4x8; do
  PPCA_B200_SOLVE64=$mode timeout 600 python bench.py --workload c3s --steps 5 --no-cpu --no-blocks > $O/bench_c3s_$mode.json 2> $O/bench_c3s_$mode.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/bench_c3s_$mode.json").read().strip().splitlines()[-1])
    print("$mode c3s ms/step", round(j["ms_per_step"],2), "solve", round(j["roofline"]["family_ms_per_step"]["solve"],2))
except Exception as e:
    print("$mode failed", e); print(open("$O/bench_c3s_$mode.err").read()[-800:])
PY
done
