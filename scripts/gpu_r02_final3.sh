#!/bin/bash
# round-2 closing measurements on one GPU: full GPU suite, the bench line (hash-tied roofline, full solve), solve times
# with the loop as one graph vs bursts, the batch benchmark and the work-precision tables
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02p_gputests.log 2>&1
echo "gpu tests exit $?"; tail -n 3 gpurun_out/r02p_gputests.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 --e2e-solve > gpurun_out/r02p_bench.json 2> gpurun_out/r02p_bench.err
echo "bench exit $?"; python -c "
import json
j=json.loads([l for l in open('gpurun_out/r02p_bench.json') if l.startswith('{')][-1])
print(j['value'], j['e2e']['value'], j['roofline'], j['e2e_solve']['value'], j['cpu_baseline']['value'])"
timeout 300 python scripts/time_solves.py > gpurun_out/r02p_time_solves.log 2>&1
echo "time solves exit $?"; cut -c1-250 gpurun_out/r02p_time_solves.log
timeout 300 python scripts/bench_batch.py --log2n 10 --batches 1,8,32 > gpurun_out/r02p_batch.log 2>&1
echo "batch exit $?"; cut -c1-400 gpurun_out/r02p_batch.log
timeout 600 python scripts/work_precision.py --out gpurun_out --tag r02 --exp-step 3 --setups fhn,henonheiles,logistic --orders 1,2,3 > gpurun_out/r02p_wp.log 2>&1
echo "work precision exit $?"; grep wrote gpurun_out/r02p_wp.log
du -sh gpurun_out
