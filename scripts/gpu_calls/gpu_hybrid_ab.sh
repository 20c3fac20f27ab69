#!/bin/bash
# A/B of the filter sweep: hybrid (Kogge-Stone top levels, default) against the plain up/down sweep (flag 16)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02s_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 12 gpurun_out/r02s_gputests.log | cut -c1-300
timeout 600 python scripts/sweep_n.py --exps 6,10,12,14,16,18,19,20 --tag r02s_sweep_hybrid > gpurun_out/r02s_sweep_hybrid.log 2>&1
echo "sweep hybrid exit $?"; cut -c1-330 gpurun_out/r02s_sweep_hybrid.log | tail -n 12
timeout 600 python scripts/sweep_n.py --exps 6,10,12,14,16,18,19,20 --flags 16 --tag r02s_sweep_updown > gpurun_out/r02s_sweep_updown.log 2>&1
echo "sweep updown exit $?"; cut -c1-330 gpurun_out/r02s_sweep_updown.log | tail -n 12
