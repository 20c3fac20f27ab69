#!/bin/bash
# First B200 run of the large-state ("tile") kernels -- they were built and host-simulated after the
# round-1 GPU budget was spent.  Run through gpurun on ONE GPU, e.g.
#   gpurun --timeout 1800 -- 'bash scripts/gpu_first_run_tile.sh r02'
# Order: correctness first (the parity tests, both Householder sweep implementations are in there), then timings at
# growing N for BOTH sweep modes with a hard timeout each (a hung kernel must not take the box), then the launch list
# and ONE ncu --set full capture of the leaf kernels at a small N per mode.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tile.py -m gpu -q > gpurun_out/${TAG}_tile_tests.log 2>&1
echo "tile tests exit $?"; tail -n 5 gpurun_out/${TAG}_tile_tests.log
for mode in smem reg; do
  export POF_B200_TILE_SWEEP=$mode
  for e in 12 15 18; do
    timeout 300 python scripts/bench_config5.py --log2n $e --steps 2 --warmup 1 \
        > gpurun_out/${TAG}_config5_${mode}_n$e.json 2> gpurun_out/${TAG}_config5_${mode}_n$e.err
    echo "config5 $mode 2^$e exit $?"; tail -c 700 gpurun_out/${TAG}_config5_${mode}_n$e.json; echo
  done
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
      --log-file gpurun_out/${TAG}_launches_config5_${mode}.csv python scripts/bench_config5.py --log2n 14 --steps 1 \
      --warmup 1 > /dev/null 2> gpurun_out/${TAG}_launches_config5_${mode}.err
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_tile_(scan|fold|smooth)' -c 3 \
      -o gpurun_out/${TAG}_prof_tile_${mode} -f python scripts/bench_config5.py --log2n 13 --steps 1 --warmup 1 \
      > gpurun_out/ncu_tile_${mode}.log 2>&1
  tail -n 2 gpurun_out/ncu_tile_${mode}.log
done
du -sh gpurun_out
