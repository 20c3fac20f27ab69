#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sweep_n.py --exps 16,19,20 --tag r02h_sweep_tma > gpurun_out/r02h_sweep_tma.log 2>&1
echo "sweep tma exit $?"; cat gpurun_out/r02h_sweep_tma.log | cut -c1-400
timeout 600 python scripts/sweep_n.py --exps 16,19,20 --flags 8 --tag r02h_sweep_notma > gpurun_out/r02h_sweep_notma.log 2>&1
echo "sweep no-tma exit $?"; cat gpurun_out/r02h_sweep_notma.log | cut -c1-400
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_large.py -m gpu -x -q > gpurun_out/r02h_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 6 gpurun_out/r02h_gputests.log | cut -c1-600
