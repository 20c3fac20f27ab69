#!/bin/bash
# round-2 third GPU call: element-form smoother suffix scan (hidden behind the filter scan), fused exchange kernels with
# virtual ranks, N sweep, bench line with parity + cpu baseline at the real N, GPU library baseline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02c_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 6 gpurun_out/r02c_gputests.log | cut -c1-600
timeout 600 python scripts/sweep_n.py --exps 6,8,10,12,14,16,18,19,20 --tag r02c_sweep_flow > gpurun_out/r02c_sweep_flow.log 2>&1
echo "sweep flow exit $?"; cat gpurun_out/r02c_sweep_flow.log | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
echo "bench exit $?"; tail -c 1800 gpurun_out/r02c_bench.json; tail -n 5 gpurun_out/r02c_bench.err
timeout 600 python bench.py --impl gpu_library --log2n 17 --steps 3 --warmup 1 > gpurun_out/r02c_gpulib_n17.json 2> gpurun_out/r02c_gpulib_n17.err
echo "gpulib 2^17 exit $?"; cat gpurun_out/r02c_gpulib_n17.json | cut -c1-700; tail -n 3 gpurun_out/r02c_gpulib_n17.err
timeout 900 python bench.py --impl gpu_library --log2n 20 --steps 2 --warmup 1 > gpurun_out/r02c_gpulib_n20.json 2> gpurun_out/r02c_gpulib_n20.err
echo "gpulib 2^20 exit $?"; cat gpurun_out/r02c_gpulib_n20.json | cut -c1-700; tail -n 3 gpurun_out/r02c_gpulib_n20.err
