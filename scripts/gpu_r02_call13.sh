#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02l_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 5 gpurun_out/r02l_gputests.log | cut -c1-400
timeout 300 python scripts/time_solves.py > gpurun_out/r02l_time_solves.log 2>&1
echo "time solves exit $?"; cat gpurun_out/r02l_time_solves.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-solve > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err
echo "bench exit $?"; python -c "
import json
j=json.loads([l for l in open('gpurun_out/r02l_bench.json') if l.startswith('{')][-1]); print(j['value'], j['e2e']['value'], j['e2e_solve'], j['roofline']['frac'])"; tail -n 3 gpurun_out/r02l_bench.err
