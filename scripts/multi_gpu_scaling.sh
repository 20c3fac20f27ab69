set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_mg.log 2>&1; tail -3 gpurun_out/pytest_mg.log
for n in 8 4 2; do
  $TR --nproc-per-node $n --master-port $((29500+n)) bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_weak_g$n.json 2> gpurun_out/scale_weak_g$n.err
  tail -c 600 gpurun_out/scale_weak_g$n.json | head -c 300; echo
done
for n in 8 4 2; do
  $TR --nproc-per-node $n --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline --n-time $((4194304/n)) > gpurun_out/scale_cfg4_g$n.json 2> gpurun_out/scale_cfg4_g$n.err
done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --n-time 4194304 > gpurun_out/scale_cfg4_g1.json 2> gpurun_out/scale_cfg4_g1.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/scale_*.json")):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1]); print(f, j["n_gpus"], j["config"]["n_time_total"], round(j["value"],4), j["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
PY
