#!/bin/bash
# round-2 GPU call N: ncu --set full of the register-tiled solve kernel at the c3 shape (source-level stall sampling)
O=gpurun_out/r02n; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:solve_tile -s 2 -c 1 -o $O/solve_tile64 -f \
  python bench.py --workload c3s --rows 131072 --steps 1 --warmup 1 --no-cpu --no-blocks > $O/ncu.log 2>&1
tail -3 $O/ncu.log
ls -la $O
