#!/bin/bash
# 1 GPU: smoother with bulk-copy (TMA) staging -- parity, A/B timing; pipelined e2e; full suite
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02g_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 6 gpurun_out/r02g_gputests.log | cut -c1-600
timeout 600 python scripts/sweep_n.py --exps 16,18,19,20 --tag r02g_sweep_tma > gpurun_out/r02g_sweep_tma.log 2>&1
echo "sweep tma exit $?"; cat gpurun_out/r02g_sweep_tma.log | cut -c1-400
timeout 600 python scripts/sweep_n.py --exps 16,18,19,20 --flags 8 --tag r02g_sweep_notma > gpurun_out/r02g_sweep_notma.log 2>&1
echo "sweep no-tma exit $?"; cat gpurun_out/r02g_sweep_notma.log | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
echo "bench exit $?"; python -c "
import json
j=json.loads([l for l in open('gpurun_out/r02g_bench.json') if l.startswith('{')][-1]); print(j['value'], j['e2e'], j['roofline'], j['cpu_baseline'], j['parity']['ok'])"; tail -n 3 gpurun_out/r02g_bench.err
