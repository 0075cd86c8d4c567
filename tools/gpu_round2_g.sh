#!/bin/bash
# round-2 GPU call G: tests, then the evidence captures: launch lists (c2 + c3 shard) and ncu --set full of the top kernels
mkdir -p gpurun_out/r02g
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02g/pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02g/pytest.log
tail -4 gpurun_out/r02g/pytest.log
# launch list of the bench command (cold-cache, serialised per-launch times; shares must agree with the CUDA-event spans)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02g/launches_c2.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-blocks > gpurun_out/r02g/launches_c2.log 2>&1
echo "launch list c2 exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02g/launches_c3s.csv \
  python bench.py --workload c3s --steps 2 --warmup 3 --no-cpu --no-blocks > gpurun_out/r02g/launches_c3s.log 2>&1
echo "launch list c3s exit $?"
# full captures: one EM step of c2 (skip the warm-up launches) and of the c3 shard
timeout 1500 ncu --set full --clock-control none --import-source on --launch-skip 60 -c 24 -o gpurun_out/r02g/c2_step \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-blocks > gpurun_out/r02g/ncu_c2.log 2>&1
echo "ncu c2 exit $?"
timeout 1500 ncu --set full --clock-control none --import-source on --launch-skip 60 -c 24 -o gpurun_out/r02g/c3s_step \
  python bench.py --workload c3s --steps 2 --warmup 3 --no-cpu --no-blocks > gpurun_out/r02g/ncu_c3s.log 2>&1
echo "ncu c3s exit $?"
for f in c2_step c3s_step; do
  ncu -i gpurun_out/r02g/$f.ncu-rep --page raw --csv > gpurun_out/r02g/${f}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02g/$f.ncu-rep --page details > gpurun_out/r02g/${f}_details.txt 2>/dev/null
  ls -la gpurun_out/r02g/$f.ncu-rep
  rm -f gpurun_out/r02g/$f.ncu-rep
done
