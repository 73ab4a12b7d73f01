# 4 GPUs: the driver's bench command (headline + parity preflight + large_grid)
N=4
mkdir -p gpurun_out
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r3i_bench_$N.json 2> gpurun_out/r3i_bench_$N.err ) 2>&1 | tail -3; echo "bench rc=$?"; tail -3 gpurun_out/r3i_bench_$N.err
python -c "
import json;d=json.load(open('gpurun_out/r3i_bench_$N.json'));print('N=$N',d['ms_per_step'],d['value'],d['parity']['ok']);
lg=d.get('large_grid') or {}
for k,v in lg.items(): print(k, v.get('ms_per_step'), v.get('value'), v.get('exchange_mode'), v.get('hbm',{}).get('frac_of_measured_peak'), v.get('nvlink'))
print(d['e2e'])"
