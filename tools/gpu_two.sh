mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "dma or p2p" 2>&1 | tail -3
for cfg in "4 4" "1 4" "4 1"; do
set -- $cfg
MLV_FWD_CHUNKS=$1 MLV_COPY_STREAMS=$2 MLV_EXCHANGE=dma timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29800 bench.py --gpus 2 --steps 10 --warmup 3 --nx 16384 --nz 16384 > gpurun_out/scale16k_dma_2.json 2> gpurun_out/scale16k_dma_2.err || tail -5 gpurun_out/scale16k_dma_2.err
python -c "
import json
d=json.loads(open('gpurun_out/scale16k_dma_2.json').read()); print('chunks $1 streams $2', d['n_gpus'], round(d['ms_per_step'],4), '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])"
done
