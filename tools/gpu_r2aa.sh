#!/bin/bash
# round-2 GPU call AA: k <= 32 on 2 x 8 register tiles (two warps per sample, three CTAs per SM)
O=gpurun_out/r02aa; mkdir -p $O
PPCA_B200_SOLVE=tile2x8 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "test_infer or test_llks or iterate_trajectory or test_mixture or smooth_extrapolate or tiled_solve" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -4 $O/pytest.log
for mode in tile2x8 default; do
  PPCA_B200_SOLVE=$mode timeout 600 python bench.py --workload c4 --rows 131072 --steps 3 --no-cpu --no-blocks > $O/bench_c4_$mode.json 2> $O/bench_c4_$mode.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/bench_c4_$mode.json").read().strip().splitlines()[-1])
    print("$mode c4 ms/step", round(j["ms_per_step"],2), "solve", round(j["roofline"]["family_ms_per_step"]["solve"],2), j.get("kernel_variants"))
except Exception as e:
    print("$mode failed", e); print(open("$O/bench_c4_$mode.err").read()[-800:])
PY
done
