mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/kbench.py 4096 4096 20 2>&1 | head -3
timeout 900 python bench.py --steps 10 --warmup 3 --nx 16384 --nz 16384 --no-cpu-baseline > gpurun_out/scale16k_1.json 2> gpurun_out/scale16k_1.err || tail -5 gpurun_out/scale16k_1.err
timeout 900 python bench.py --steps 20 --warmup 3 --nx 8192 --nz 8192 --no-cpu-baseline > gpurun_out/bench8k_1.json 2> gpurun_out/bench8k_1.err || tail -5 gpurun_out/bench8k_1.err
for f in scale16k_1 bench8k_1; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read()); print('$f', d['n_gpus'], round(d['ms_per_step'],4), '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['step'], {k:v['ms'] for k,v in d['roofline']['kernels'].items()})"; done
