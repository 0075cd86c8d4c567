#!/bin/bash
# round-2 GPU call D (2 GPUs): two-GPU tests of the sharded C ABI, torchrun bench at N=2 (native NCCL path), c4 probes
mkdir -p gpurun_out/r02d
nvidia-smi -L > gpurun_out/r02d/gpus.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_gpus or two_contexts or world_of_one" > gpurun_out/r02d/pytest_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d/pytest_2gpu.log
tail -4 gpurun_out/r02d/pytest_2gpu.log
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02d/bench_c2_2gpu.json 2> gpurun_out/r02d/bench_c2_2gpu.err; echo "bench2 exit $?"
tail -3 gpurun_out/r02d/bench_c2_2gpu.err
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02d/bench_c2_2gpu.json").read().strip().splitlines()[-1])
print("c2 x2", j["ms_per_step"], j["value"], "e2e", j["e2e"]["value"], j["e2e"]["h2d_gb_per_s"])
print("strong", j["strong_scaling"])
for b in ("c3_shard","c4_shard"):
    x=j[b]; print(b, x["ms_per_step"], x["value"], "comm ms", x["comm_ms_per_step"], x.get("comm_gb_per_s"))
PY
for g in 1 0; do
PPCA_B200_GUARD=$g timeout 600 python bench.py --workload c4 --rows 131072 --steps 3 --no-cpu --no-blocks > gpurun_out/r02d/bench_c4_probe_guard$g.json 2> gpurun_out/r02d/bench_c4_probe_guard$g.err
python - <<PY
import json
j=json.loads(open("gpurun_out/r02d/bench_c4_probe_guard$g.json").read().strip().splitlines()[-1])
print("c4 probe guard=$g", j["ms_per_step"], j["roofline"]["family_ms_per_step"], j["kernel_variants"], j["gpu_launches"])
PY
done
