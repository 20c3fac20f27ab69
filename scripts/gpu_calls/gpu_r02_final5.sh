#!/bin/bash
# last call of the round: GPU suite with the smoother's hybrid suffix scan, smoke, short N sweep
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r02w_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 3 gpurun_out/r02w_gputests.log | cut -c1-300
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2 | cut -c1-300
timeout 100 python scripts/sweep_n.py --exps 10,12,14,16,18,20 --tag r02w_sweep > gpurun_out/r02w_sweep.log 2>&1
echo "sweep exit $?"; cut -c1-330 gpurun_out/r02w_sweep.log | tail -n 8
