seg() { python -c "
import json,sys; j=json.load(open(sys.argv[1])); print(sys.argv[1], round(j['value'],4), round(j['e2e']['value'],3), j['gpu_launches'], {k:round(v,4) for k,v in j['roofline_iteration']['segments_ms'].items()})" $1; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest19.log 2>&1; tail -4 gpurun_out/pytest19.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b_A.json 2>gpurun_out/b_A.err; seg gpurun_out/b_A.json; tail -3 gpurun_out/b_A.err
