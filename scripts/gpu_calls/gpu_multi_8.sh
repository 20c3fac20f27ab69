#!/bin/bash
# 8 GPUs: NCCL + peer-memory parity tests, weak and strong scaling (config 4) at 8 ranks
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "nccl_matches or solve_sharded_nccl" > gpurun_out/r02j_sharded_tests.log 2>&1
echo "sharded tests exit $?"; tail -n 4 gpurun_out/r02j_sharded_tests.log | cut -c1-400
for sc in weak strong; do
  for ex in p2p nccl; do
    POF_B200_EXCHANGE=$ex timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 \
      bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline --scaling $sc > gpurun_out/r02j_${sc}_${ex}_g8.json 2> gpurun_out/r02j_${sc}_${ex}_g8.err
    echo "bench $sc $ex exit $?"; python - <<PY
import json
try:
    j=json.loads([x for x in open("gpurun_out/r02j_${sc}_${ex}_g8.json") if x.startswith("{")][-1])
    print(j["value"], j["e2e"]["value"], j["e2e"]["serial_value"], j["config"]["exchange"], j["parity"]["ok"], j["config"]["n_time_total"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/r02j_${sc}_${ex}_g8.err").read()[-1500:])
PY
  done
done
