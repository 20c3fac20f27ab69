#!/bin/bash
# round-2 final GPU call (1 GPU): full GPU suite, smoke, bench line, ncu refresh tied to the final kernel sources,
# work-precision sweep (config 3) with the final kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 5 gpurun_out/r02k_gputests.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02k_smoke.log 2>&1
echo "smoke exit $?"; tail -n 2 gpurun_out/r02k_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
echo "bench exit $?"; python -c "
import json
j=json.loads([l for l in open('gpurun_out/r02k_bench.json') if l.startswith('{')][-1]); print(j['value'], j['e2e'], j['roofline']['frac'], j['cpu_baseline']['value'], j['parity']['ok'])"; tail -n 3 gpurun_out/r02k_bench.err
bash scripts/gpu_profiles.sh r02 > gpurun_out/r02k_profiles.log 2>&1
echo "profiles exit $?"; tail -n 6 gpurun_out/r02k_profiles.log
timeout 1500 python scripts/work_precision.py --out gpurun_out --tag r02 --exp-step 3 > gpurun_out/r02k_wp.log 2>&1
echo "work precision exit $?"; grep wrote gpurun_out/r02k_wp.log
