#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_loop.py -x -q > gpurun_out/r02m_looptests.log 2>&1
echo "loop tests exit $?"; tail -n 25 gpurun_out/r02m_looptests.log | cut -c1-300
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02m_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 3 gpurun_out/r02m_gputests.log | cut -c1-300
timeout 300 python scripts/time_solves.py > gpurun_out/r02m_time_solves.log 2>&1
echo "time solves exit $?"; cat gpurun_out/r02m_time_solves.log | cut -c1-300
