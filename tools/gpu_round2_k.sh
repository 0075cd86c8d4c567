#!/bin/bash
# round-2 GPU call K: compute-sanitizer over the round-2 kernels (memcheck: everything; racecheck: solve / contraction / mixture)
mkdir -p gpurun_out/r02k
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/r02k/memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/r02k/memcheck.log
tail -5 gpurun_out/r02k/memcheck.log
SANITIZE_ONLY=0,2,3,x SANITIZE_MIX=0 timeout 2400 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/r02k/racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/r02k/racecheck.log
tail -5 gpurun_out/r02k/racecheck.log
