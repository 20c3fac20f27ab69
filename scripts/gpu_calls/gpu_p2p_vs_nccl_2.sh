#!/bin/bash
# 2 GPUs: peer-memory exchange kernels vs NCCL all-gathers (tests + A/B bench lines)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "nccl_matches or solve_sharded_nccl" > gpurun_out/r02f_sharded_tests.log 2>&1
echo "sharded tests exit $?"; tail -n 8 gpurun_out/r02f_sharded_tests.log | cut -c1-500
for ex in nccl p2p; do
  for sc in weak strong; do
    POF_B200_EXCHANGE=$ex timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline --scaling $sc > gpurun_out/r02f_${sc}_${ex}_g2.json 2> gpurun_out/r02f_${sc}_${ex}_g2.err
    echo "bench $sc $ex exit $?"; python - <<PY
import json
try:
    j=json.loads([x for x in open("gpurun_out/r02f_${sc}_${ex}_g2.json") if x.startswith("{")][-1])
    print(j["value"], j["e2e"]["value"], j["config"]["exchange"], j["parity"]["ok"], j["gpu_launches_per_step"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/r02f_${sc}_${ex}_g2.err").read()[-1500:])
PY
  done
done
