#!/bin/bash
# round-2 GPU call O: the widened rows (f1 full covariances, f3 device samplers, f4 device ingestion) + solve tests
O=gpurun_out/r02o; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "from_device or samplers or full_covariances or state_size_sweep or covariance_diagonals or posterior_samplers_carry" > $O/pytest_new.log 2>&1; echo "pytest exit $?" >> $O/pytest_new.log
tail -25 $O/pytest_new.log
