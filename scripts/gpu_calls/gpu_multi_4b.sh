#!/bin/bash
# 4 GPUs: end-to-end (host-buffer) numbers of the weak-scaling line with each rank bound to the CPUs next to its GPU
mkdir -p gpurun_out
nvidia-smi topo -m 2>/dev/null | head -12 > gpurun_out/r02q_topo.txt
run() { # gpus, tag, extra args
  g=$1; tag=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + g)) \
      bench.py --gpus $g --steps 30 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r02q_${tag}_g$g.json 2> gpurun_out/r02q_${tag}_g$g.err
  echo "bench $tag g=$g exit $?"; python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r02q_${tag}_g$g.json") if x.startswith("{")][-1]; j=json.loads(l)
    print(j["value"], "e2e", j["e2e"]["value"], j["e2e"]["serial_value"], j["config"]["cpu_affinity"], j["config"]["exchange"], j["parity"]["ok"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/r02q_${tag}_g$g.err").read()[-1500:])
PY
}
run 2 weak
run 4 weak
run 4 strong --scaling strong
