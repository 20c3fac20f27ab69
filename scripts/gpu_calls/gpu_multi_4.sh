#!/bin/bash
# round-2 fourth GPU call (4 GPUs): NCCL parity tests, weak / strong scaling lines with the fused exchange kernels
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_userf.py -m gpu -x -q > gpurun_out/r02d_sharded_tests.log 2>&1
echo "sharded tests exit $?"; tail -n 8 gpurun_out/r02d_sharded_tests.log | cut -c1-500
run() { # gpus, tag, extra args
  g=$1; tag=$2; shift 2
  if [ "$g" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r02d_${tag}_g$g.json 2> gpurun_out/r02d_${tag}_g$g.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port $((29500 + g)) \
      bench.py --gpus $g --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/r02d_${tag}_g$g.json 2> gpurun_out/r02d_${tag}_g$g.err
  fi
  echo "bench $tag g=$g exit $?"; python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/r02d_${tag}_g$g.json") if x.startswith("{")][-1]; j=json.loads(l)
    print({k:j[k] for k in ("value","n_gpus","scaling","gpu_launches_per_step")}, j["e2e"]["value"], j["parity"], j["config"]["n_time_total"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/r02d_${tag}_g$g.err").read()[-1500:])
PY
}
run 1 weak
run 2 weak
run 4 weak
run 1 strong --scaling strong
run 2 strong --scaling strong
run 4 strong --scaling strong
