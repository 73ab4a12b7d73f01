# last evidence of round 2: the driver's bench command on the final sources (roofline.traffic filled from the
# digest-checked ncu captures), the config-3 line, then the GPU parity suite with whatever time is left
mkdir -p gpurun_out
( time timeout 130 python bench.py --steps 200 --warmup 10 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err ) 2>&1 | grep real; echo "bench rc=$?"
timeout 40 python bench.py --config rbc --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02d_bench_rbc.json 2> gpurun_out/r02d_bench_rbc.err; echo "rbc rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r02d_bench.json')); print(d['ms_per_step'], d['roofline']['step']['frac'], d['roofline']['traffic'], d['e2e']['ms_per_step'], d['parity']['ok'])
for k,v in (d.get('large_grid') or {}).items(): print(k, v['ms_per_step'], v['hbm']['frac_of_measured_peak'])
r=json.load(open('gpurun_out/r02d_bench_rbc.json')); print('rbc', r['ms_per_step'], r['roofline']['frac'], r['roofline']['traffic'])"
timeout 70 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02d_pytest_gpu.log
