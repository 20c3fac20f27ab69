timeout 400 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_mg2.log 2>&1; tail -4 gpurun_out/pytest_mg2.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/b_g2.json 2> gpurun_out/b_g2.err; python -c "
import json; j=json.load(open('gpurun_out/b_g2.json')); print(j['n_gpus'], j['value'], j['e2e']['value'])"; tail -2 gpurun_out/b_g2.err
