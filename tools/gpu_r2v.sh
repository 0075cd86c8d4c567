#!/bin/bash
# round-2 GPU call V: k <= 16 through the register-tiled, panel-blocked solve (8 threads per sample, four samples per warp)
O=gpurun_out/r02v; mkdir -p $O
PPCA_B200_SOLVE16=tile timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "test_llks_and_llk or test_infer or test_iterate_trajectory or golden_single or test_priors or precision_guard or host_streaming or test_mixture" > $O/pytest_tile16.log 2>&1; echo "pytest exit $?" >> $O/pytest_tile16.log
tail -6 $O/pytest_tile16.log
for mode in tile rows; do
  PPCA_B200_SOLVE16=$mode timeout 600 python bench.py --no-cpu --no-blocks > $O/bench_c2_$mode.json 2> $O/bench_c2_$mode.err
  python - <<PY
import json
try:
    j=json.loads(open("$O/bench_c2_$mode.json").read().strip().splitlines()[-1])
    print("$mode c2 ms/step", round(j["ms_per_step"],3), "solve", round(j["roofline"]["family_ms_per_step"]["solve"],3), j.get("kernel_variants"))
except Exception as e:
    print("$mode failed", e); print(open("$O/bench_c2_$mode.err").read()[-800:])
PY
done
