python -m pytest tests/test_gpu_rrs.py -x -q -m gpu 2>&1 | tail -5
python bench.py --config 5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_cfg5_v2.json 2>gpurun_out/err.txt; tail -3 gpurun_out/err.txt
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_cfg5_v2.json')); print(d['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['roofline']['elastic_export_ms'], d['parity_check'])"
