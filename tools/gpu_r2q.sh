# N GPUs: sharded parity tests (all exchange modes, sharding behind the public API) + the driver's bench command
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv | tail -n +2 | sort | uniq -c
timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r2q_sharded_tests_$N.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2q_sharded_tests_$N.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2q_bench_$N.json 2> gpurun_out/r2q_bench_$N.err ) 2>&1 | tail -3; echo "bench rc=$?"; tail -5 gpurun_out/r2q_bench_$N.err
python -c "
import json;d=json.load(open('gpurun_out/r2q_bench_$N.json'));print('N=$N',d['ms_per_step'],d['value'],d['parity'],d.get('config',{}).get('parallelism'));
lg=d.get('large_grid') or {}
for k,v in lg.items(): print(k, v.get('ms_per_step'), v.get('value'), v.get('exchange_mode'), v.get('hbm',{}).get('frac_of_measured_peak'))
print(d['e2e'])"
