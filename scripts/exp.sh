seg() { python -c "
import json,sys; j=json.load(open(sys.argv[1])); print(sys.argv[1], round(j['value'],4), {k:round(v,4) for k,v in j['roofline_iteration']['segments_ms'].items()})" $1; }
python -m pytest tests -m gpu -x -q > gpurun_out/pytest16.log 2>&1; tail -2 gpurun_out/pytest16.log
python bench.py --no-cpu-baseline > gpurun_out/b_A.json 2>gpurun_out/b_A.err; seg gpurun_out/b_A.json
POF_B200_LIB=$PWD/parallel-in-time-ode-filters_b200/variants/libpof_mz2.so python bench.py --no-cpu-baseline > gpurun_out/b_B.json 2>gpurun_out/b_B.err; seg gpurun_out/b_B.json
