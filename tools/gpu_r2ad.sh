#!/bin/bash
# round-2 GPU call AD: tiled solve with one shared array of column maxima and cp.async prefetch of the next sample's packed G
O=gpurun_out/r02ad; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "state_size_sweep or tiled_solve or two_tile or test_infer or test_llks or iterate_trajectory or test_mixture or smooth_extrapolate or host_streamed" > $O/pytest.log 2>&1; echo "pytest exit $?" >> $O/pytest.log
tail -4 $O/pytest.log
timeout 600 python bench.py --workload c3s --steps 5 --no-cpu --no-blocks > $O/bench_c3s.json 2> $O/bench_c3s.err
timeout 600 python bench.py --workload c5 --steps 3 --no-cpu --no-blocks > $O/bench_c5.json 2> $O/bench_c5.err
python - <<PY
import json
for wl in ("c3s","c5"):
    try:
        j=json.loads(open("$O/bench_%s.json"%wl).read().strip().splitlines()[-1])
        print(wl, "ms/step", round(j["ms_per_step"],2), "solve", round(j["roofline"]["family_ms_per_step"]["solve"],2))
    except Exception as e:
        print(wl, "failed", e); print(open("$O/bench_%s.err"%wl).read()[-600:])
PY
