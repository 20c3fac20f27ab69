#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/sweep_n.py --exps 19,20 --tag r02i_sweep_tma > gpurun_out/r02i_sweep_tma.log 2>&1
echo "sweep tma exit $?"; cat gpurun_out/r02i_sweep_tma.log | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lane2 and not no_tma" > gpurun_out/r02i_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 3 gpurun_out/r02i_gputests.log | cut -c1-300
