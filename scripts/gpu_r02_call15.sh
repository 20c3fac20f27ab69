#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_loop.py -x -q 2>&1 | tail -n 3 | cut -c1-300
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02n_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 3 gpurun_out/r02n_gputests.log | cut -c1-300
bash scripts/gpu_profiles.sh r02
timeout 900 python bench.py --steps 20 --warmup 5 --e2e-solve > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
echo "bench exit $?"; tail -c 600 gpurun_out/r02n_bench.json; tail -n 3 gpurun_out/r02n_bench.err
