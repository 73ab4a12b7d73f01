# large-grid kernels after the specialisations + full bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_abi.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
echo "== 16384^2"; timeout 600 python tools/kbench.py 16384 16384 10 2>&1 | head -5
echo "== 8192^2"; timeout 600 python tools/kbench.py 8192 8192 20 2>&1 | head -5
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2y_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2y_bench.json'));print(d['ms_per_step'],d['value'],d['roofline']['step']['frac']); print({k:v['ms'] for k,v in d['roofline']['kernels'].items()}); print(d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['blocking']['ms_per_step'])
for k,v in (d.get('large_grid') or {}).items(): print(k, v['ms_per_step'], v['hbm']['frac_of_measured_peak'])"
