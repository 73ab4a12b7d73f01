# full GPU suite + bench after: optional reductions, trig bases, ensemble
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2s_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3
timeout 900 python bench.py --steps 100 --warmup 10 --no-large-grid --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2s_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2s_bench.json'));print(d['ms_per_step'],d['value'],d['roofline']['step']['frac']); print({k:v['ms'] for k,v in d['roofline']['kernels'].items()}); print(d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['blocking']['ms_per_step'])"
