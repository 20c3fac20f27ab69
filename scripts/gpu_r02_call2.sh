#!/bin/bash
# round-2 second GPU call: the dataflow tree sweeps -- correctness (full GPU suite), then A/B against one launch per level
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02b_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 12 gpurun_out/r02b_gputests.log | cut -c1-400
grep "N=2^" gpurun_out/r02b_gputests.log | sort -u
timeout 600 python scripts/sweep_n.py --exps 6,8,10,12,14,16,18,19,20 --tag r02b_sweep_flow > gpurun_out/r02b_sweep_flow.log 2>&1
echo "sweep flow exit $?"; cat gpurun_out/r02b_sweep_flow.log | cut -c1-400
timeout 600 python scripts/sweep_n.py --exps 6,10,14,20 --flags 4 --tag r02b_sweep_perlevel > gpurun_out/r02b_sweep_perlevel.log 2>&1
echo "sweep per-level exit $?"; cat gpurun_out/r02b_sweep_perlevel.log | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "bench exit $?"; tail -c 3000 gpurun_out/r02b_bench.json; tail -n 5 gpurun_out/r02b_bench.err
