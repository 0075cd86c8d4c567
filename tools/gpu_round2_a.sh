#!/bin/bash
# round-2 GPU call A: full GPU test suite, compute-sanitizer passes, c2 bench line
mkdir -p gpurun_out/r02a
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02a/gpu.txt 2>&1
nproc >> gpurun_out/r02a/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02a/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02a/pytest.log
tail -5 gpurun_out/r02a/pytest.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/r02a/memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02a/memcheck.log
tail -4 gpurun_out/r02a/memcheck.log
SANITIZE_ONLY=0,2,5 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/r02a/racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r02a/racecheck.log
tail -4 gpurun_out/r02a/racecheck.log
timeout 600 python bench.py > gpurun_out/r02a/bench_c2.json 2> gpurun_out/r02a/bench_c2.err; echo "bench exit $?"
tail -c 1500 gpurun_out/r02a/bench_c2.json
