#!/bin/bash
# closing measurements with the hybrid filter sweep: GPU suite, launch list + ncu captures, bench line, N sweep
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02u_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 3 gpurun_out/r02u_gputests.log | cut -c1-300
bash scripts/gpu_profiles.sh r02 > gpurun_out/r02u_profiles.log 2>&1
echo "profiles exit $?"; tail -n 3 gpurun_out/r02u_profiles.log
timeout 900 python bench.py --steps 20 --warmup 5 --e2e-solve > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err
echo "bench exit $?"; python -c "
import json
j=json.loads([l for l in open('gpurun_out/r02u_bench.json') if l.startswith('{')][-1])
print(j['value'], j['e2e']['value'], j['roofline']['frac'], j['e2e_solve']['value'], j['cpu_baseline']['value'], j['gpu_launches_per_step'])"
timeout 600 python scripts/sweep_n.py --tag r02u_sweep > gpurun_out/r02u_sweep.log 2>&1
echo "sweep exit $?"; cut -c1-120 gpurun_out/r02u_sweep.log | tail -n 16
timeout 300 python scripts/time_solves.py > gpurun_out/r02u_time_solves.log 2>&1
echo "time solves exit $?"; grep while-graph gpurun_out/r02u_time_solves.log | cut -c1-200
du -sh gpurun_out
