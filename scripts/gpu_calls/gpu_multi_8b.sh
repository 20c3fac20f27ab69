#!/bin/bash
# 8 GPUs: weak and strong scaling lines (peer-memory exchange), ranks bound to the CPUs next to their GPU
mkdir -p gpurun_out
for sc in weak strong; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29548 \
      bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline --scaling $sc > gpurun_out/r02q_${sc}_g8.json 2> gpurun_out/r02q_${sc}_g8.err
  echo "bench $sc exit $?"; python - <<PY
import json
try:
    j=json.loads([x for x in open("gpurun_out/r02q_${sc}_g8.json") if x.startswith("{")][-1])
    print(j["value"], "e2e", j["e2e"]["value"], j["e2e"]["serial_value"], j["config"]["cpu_affinity"], j["config"]["exchange"], j["parity"]["ok"], j["config"]["n_time_total"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/r02q_${sc}_g8.err").read()[-1500:])
PY
done
