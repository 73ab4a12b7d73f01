# round 2, call g (N GPUs): exchange modes at 4096^2 (peer stores + device flags + CUDA graph, copy engines, NCCL)
N=${1:-2}
mkdir -p gpurun_out
for m in p2p dma a2a; do
MLV_EXCHANGE=$m timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 200 --warmup 10 --no-large-grid > gpurun_out/r2g_bench${N}_$m.json 2> gpurun_out/r2g_bench${N}_$m.err; echo "bench $m rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r2g_bench${N}_$m.json'));print('$m N=$N',d['ms_per_step'],d['value'],d['parity']['ok'],d['gpu_launches'])"
done
MLV_EXCHANGE=p2p MLV_GRAPH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 200 --warmup 10 --no-large-grid > gpurun_out/r2g_bench${N}_p2p_nograph.json 2> gpurun_out/r2g_bench${N}_p2p_nograph.err; python -c "
import json;d=json.load(open('gpurun_out/r2g_bench${N}_p2p_nograph.json'));print('p2p nograph N=$N',d['ms_per_step'],d['value'],d['parity']['ok'],d['gpu_launches'])"
