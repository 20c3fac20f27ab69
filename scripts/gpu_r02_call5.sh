#!/bin/bash
# round-2 fifth GPU call (1 GPU): full suite incl. virtual ranks / user f / fp32 / QPM, fp32 accuracy map, bench lines
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02e_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 8 gpurun_out/r02e_gputests.log | cut -c1-600
timeout 900 python scripts/fp32_accuracy.py > gpurun_out/r02e_fp32.log 2>&1
echo "fp32 accuracy exit $?"; tail -n 12 gpurun_out/r02e_fp32.log; cat gpurun_out/r02_fp32_accuracy.md
timeout 600 python bench.py --dtype f32 --steps 20 --warmup 5 > gpurun_out/r02e_bench_f32.json 2> gpurun_out/r02e_bench_f32.err
echo "bench f32 exit $?"; tail -c 600 gpurun_out/r02e_bench_f32.json; tail -n 3 gpurun_out/r02e_bench_f32.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --start converged > gpurun_out/r02e_bench_converged.json 2> gpurun_out/r02e_bench_converged.err
echo "bench converged exit $?"; tail -c 900 gpurun_out/r02e_bench_converged.json; tail -n 3 gpurun_out/r02e_bench_converged.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-solve > gpurun_out/r02e_bench_e2esolve.json 2> gpurun_out/r02e_bench_e2esolve.err
echo "bench e2e-solve exit $?"; python -c "
import json
j=json.loads([l for l in open('gpurun_out/r02e_bench_e2esolve.json') if l.startswith('{')][-1]); print(j['value'], j['e2e'], j['e2e_solve'])"; tail -n 3 gpurun_out/r02e_bench_e2esolve.err
