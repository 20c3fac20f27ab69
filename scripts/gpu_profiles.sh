#!/bin/bash
# GPU side of the profile refresh (one GPU, through gpurun): launch list of one bench run (kernel shares of a step),
# ncu --set full of the leaf kernels and the tree kernels; with "tile" as second argument also the large-state kernels
# of config 5 (the three reports together exceed gpurun's 64 MiB return limit: capture them in separate calls).
# The summaries under profiles/ are produced in the container by scripts/summarize_ncu.py.
TAG=${1:-r02}
mkdir -p gpurun_out
if [ "$2" != "tile" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_lane2 -c 3 -o gpurun_out/${TAG}_prof_lane2 -f \
    python scripts/profile_iter.py 1048576 1 > gpurun_out/ncu_lane2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_tree' -c 6 -o gpurun_out/${TAG}_prof_tree -f \
    python scripts/profile_iter.py 1048576 1 > gpurun_out/ncu_tree.log 2>&1
tail -n 2 gpurun_out/ncu_lane2.log; tail -n 2 gpurun_out/ncu_tree.log
else
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tile_(fold|scan|smooth)' -c 3 -o gpurun_out/${TAG}_prof_tile -f \
    python scripts/bench_config5.py --log2n 13 --steps 1 --warmup 1 > gpurun_out/ncu_tile.log 2>&1
tail -n 2 gpurun_out/ncu_tile.log
fi
du -sh gpurun_out; ls -la gpurun_out/*.ncu-rep
