#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_profiles.sh r02 > gpurun_out/r02k_profiles.log 2>&1
echo "profiles exit $?"; tail -n 4 gpurun_out/r02k_profiles.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
echo "bench exit $?"; tail -c 400 gpurun_out/r02k_bench.json
timeout 300 python scripts/bench_batch.py --log2n 10 --batches 1,8,32 > gpurun_out/r02k_batch.log 2>&1
echo "batch exit $?"; cat gpurun_out/r02k_batch.log | cut -c1-400
timeout 600 python scripts/work_precision.py --out gpurun_out --tag r02 --exp-step 3 --setups fhn,henonheiles,logistic --orders 1,2,3 > gpurun_out/r02k_wp.log 2>&1
echo "work precision exit $?"; grep wrote gpurun_out/r02k_wp.log
du -sh gpurun_out
