# N GPUs (4 or 8): exchange modes at 4096^2, then the driver's bench command (headline + large_grid)
N=${1:-8}
mkdir -p gpurun_out
for m in p2p dma; do
MLV_EXCHANGE=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 200 --warmup 10 --no-large-grid > gpurun_out/r2r_bench${N}_$m.json 2> gpurun_out/r2r_bench${N}_$m.err; echo "bench $m rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r2r_bench${N}_$m.json'));print('$m N=$N',d['ms_per_step'],d['value'],d['parity']['ok'],d['gpu_launches'])"
done
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2r_bench_$N.json 2> gpurun_out/r2r_bench_$N.err ) 2>&1 | tail -3; echo "bench rc=$?"; tail -3 gpurun_out/r2r_bench_$N.err
python -c "
import json;d=json.load(open('gpurun_out/r2r_bench_$N.json'));print('N=$N',d['ms_per_step'],d['value'],d['parity']['ok']);
lg=d.get('large_grid') or {}
for k,v in lg.items(): print(k, v.get('ms_per_step'), v.get('value'), v.get('exchange_mode'), v.get('hbm',{}).get('frac_of_measured_peak'), v.get('nvlink'))
print(d['e2e'])"
