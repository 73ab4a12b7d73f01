mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3
for mode in dma; do
MLV_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29800 bench.py --gpus 2 --steps 10 --warmup 3 --nx 16384 --nz 16384 > gpurun_out/scale16k_${mode}_2.json 2> gpurun_out/scale16k_${mode}_2.err || tail -5 gpurun_out/scale16k_${mode}_2.err
python -c "
import json
d=json.loads(open('gpurun_out/scale16k_${mode}_2.json').read()); print('$mode', d['n_gpus'], round(d['ms_per_step'],4), '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])"
done
