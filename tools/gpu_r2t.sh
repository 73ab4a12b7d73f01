# reductions compiled out on non-ticker steps: quick parity + kbench (red on) + bench (auto) + rbc
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_abi.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2t_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2t_bench.json'));print(d['ms_per_step'],d['value'],d['roofline']['step']['frac']); print({k:v['ms'] for k,v in d['roofline']['kernels'].items()}); print(d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['blocking']['ms_per_step'])
for k,v in (d.get('large_grid') or {}).items(): print(k, v['ms_per_step'], v['hbm']['frac_of_measured_peak'])"
timeout 600 python bench.py --config rbc --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2t_bench_rbc.json 2> gpurun_out/r2t_bench_rbc.err; python -c "
import json;d=json.load(open('gpurun_out/r2t_bench_rbc.json'));print('rbc',d['ms_per_step'],d['value'],d['roofline']['frac'])"
