#!/bin/bash
# GPU side of the profile refresh (run through gpurun, one GPU):
#   1. launch list of one bench run (kernel shares of a step)
#   2. ncu --set full of the three leaf kernels and of the first 16 tree launches of one iteration (the reports must
#      stay below gpurun's 64 MiB return limit)
# The summaries under profiles/ are produced in the container by scripts/summarize_ncu.py.
TAG=${1:-r01f}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
ncu --set full --clock-control none --import-source on -k regex:k_lane2 -c 3 -o gpurun_out/${TAG}_prof_lane2 -f \
    python scripts/profile_iter.py 1048576 1 > gpurun_out/ncu_lane2.log 2>&1
ncu --set full --clock-control none -k regex:k_tree -c 16 -o gpurun_out/${TAG}_prof_tree -f \
    python scripts/profile_iter.py 1048576 1 > gpurun_out/ncu_tree.log 2>&1
tail -n 2 gpurun_out/ncu_lane2.log; tail -n 2 gpurun_out/ncu_tree.log; du -sh gpurun_out
