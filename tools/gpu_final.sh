# final evidence of the round on one GPU: parity suite, smoke, bench lines, launch list, full ncu capture
TAG=${1:-r02c}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tail -1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -4 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ) 2>&1 | tail -3; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; cut -c1-200 gpurun_out/${TAG}_bench_ref.json
timeout 600 python bench.py --config rbc --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_rbc.json 2> gpurun_out/${TAG}_bench_rbc.err; echo "rbc rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_xinv|k_z_advect|k_xfwd|k_reduce" -s 62 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-large-grid > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_xinv|k_z_advect|k_xfwd" -s 74 -c 3 -o gpurun_out/prof_${TAG} -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-large-grid > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep
python -c "
import json;d=json.load(open('gpurun_out/${TAG}_bench.json'));print(d['ms_per_step'],d['value'],d['roofline']['step']['frac']); print({k:v['ms'] for k,v in d['roofline']['kernels'].items()}); print(d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['blocking']['ms_per_step'], d['cpu_baseline']['value'])
for k,v in (d.get('large_grid') or {}).items(): print(k, v['ms_per_step'], v['hbm']['frac_of_measured_peak'])
r=json.load(open('gpurun_out/${TAG}_bench_rbc.json'));print('rbc',r['ms_per_step'],r['value'],r['roofline']['frac'])"
