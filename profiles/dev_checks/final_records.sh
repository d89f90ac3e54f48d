# final records of the round on ONE B200: GPU test suite, bench lines of every config, launch list + ncu summary of config 5
set -x
python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_gputests.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_cfg2.json 2> gpurun_out/r02_bench_cfg2.err
python bench.py --config 5 --steps 3 --warmup 3 > gpurun_out/r02_bench_cfg5.json 2> gpurun_out/r02_bench_cfg5.err
python bench.py --config 3 --steps 2 --warmup 3 --cpu-seconds 10 > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/r02_bench_cfg3.err
python bench.py --impl reference --config 5 --steps 1 --warmup 0 --cpu-seconds 10 > gpurun_out/r02_bench_cfg5_reference_arm.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_cfg5.csv python bench.py --config 5 --steps 1 --warmup 0 --no-cpu-baseline --no-parity-check > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_rrs_inel -s 9 -c 1 -o gpurun_out/r02_rrs_inel_final python bench.py --config 5 --steps 1 --warmup 0 --no-cpu-baseline --no-parity-check > /dev/null 2>&1
python profiles/dev_checks/small_vs_generic.py > gpurun_out/r02_small_vs_generic.json 2>/dev/null
tail -c 400 gpurun_out/r02_gputests.log
for f in cfg2 cfg5 cfg3; do python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/r02_bench_$f.json') if l.startswith('{')][0]); print('$f', d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('parity_check',{}).get('ok'), d.get('cpu_baseline',{}).get('value'))"; done
