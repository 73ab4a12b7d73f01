mkdir -p gpurun_out
for n in 8 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29800 bench.py --gpus $n --steps 10 --warmup 3 --nx 16384 --nz 16384 > gpurun_out/scale16k_dma_$n.json 2> gpurun_out/scale16k_dma_$n.err || tail -5 gpurun_out/scale16k_dma_$n.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29800 bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/scale4k_$n.json 2> gpurun_out/scale4k_$n.err || tail -5 gpurun_out/scale4k_$n.err
for f in scale16k_dma_$n scale4k_$n; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read()); print('$f', d['config']['exchange_mode'], d['n_gpus'], round(d['ms_per_step'],4), '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])"; done
done
