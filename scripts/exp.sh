timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest22.log 2>&1; tail -3 gpurun_out/pytest22.log
timeout 600 python bench.py > gpurun_out/bench10.json 2>gpurun_out/bench10.err; cat gpurun_out/bench10.json | head -c 600; echo; tail -3 gpurun_out/bench10.err
timeout 300 python scripts/sweep_n.py --graph > gpurun_out/sweep_graph.log 2>&1; tail -15 gpurun_out/sweep_graph.log
bash scripts/capture_profiles.sh r01f
