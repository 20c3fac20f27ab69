#!/bin/bash
# round-2 first GPU call: full GPU test suite (incl. the 2^17 / 2^20 parity cases), config-5 A/B of the two tile
# Householder sweep modes, a bench line of the round-1 kernel state as this round's starting point
mkdir -p gpurun_out
nproc > gpurun_out/r02_nproc.txt; free -g >> gpurun_out/r02_nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02a_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 15 gpurun_out/r02a_gputests.log
grep "N=2^" gpurun_out/r02a_gputests.log
for mode in smem reg; do
  export POF_B200_TILE_SWEEP=$mode
  for e in 12 15 18; do
    timeout 300 python scripts/bench_config5.py --log2n $e --steps 2 --warmup 1 \
        > gpurun_out/r02a_config5_${mode}_n$e.json 2> gpurun_out/r02a_config5_${mode}_n$e.err
    echo "config5 $mode 2^$e exit $?"; tail -c 900 gpurun_out/r02a_config5_${mode}_n$e.json; echo
  done
done
unset POF_B200_TILE_SWEEP
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench exit $?"; tail -c 1500 gpurun_out/r02a_bench.json
