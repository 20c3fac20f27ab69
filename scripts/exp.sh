seg() { python -c "
import json,sys; j=json.load(open(sys.argv[1])); print(sys.argv[1], j['n_gpus'], round(j['value'],4), round(j['e2e']['value'],3), j['gpu_launches'], {k:round(v,4) for k,v in j['roofline_iteration']['segments_ms'].items()}, j['roofline_iteration']['ms_per_step_eager_launches'])" $1; }
timeout 300 python scripts/dbg_highorder.py > gpurun_out/dbg_highorder.log 2>&1; cat gpurun_out/dbg_highorder.log | grep -v Warn | tail -40
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest23.log 2>&1; tail -3 gpurun_out/pytest23.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b_A.json 2>gpurun_out/b_A.err; seg gpurun_out/b_A.json; tail -3 gpurun_out/b_A.err
POF_B200_TREE_APEX=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b_B.json 2>gpurun_out/b_B.err; seg gpurun_out/b_B.json; tail -3 gpurun_out/b_B.err
timeout 300 python scripts/sweep_n.py --graph > gpurun_out/sweep_graph.log 2>&1; tail -15 gpurun_out/sweep_graph.log | tr '\n' ' '
