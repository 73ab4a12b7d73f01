# 8-GPU session: sharded parity (incl. long-line kernels) + strong scaling of the 16384^2 grid (N = 2, 4, 8)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29800 bench.py --gpus $n --steps 10 --warmup 3 --nx 16384 --nz 16384 > gpurun_out/scale16k_$n.json 2> gpurun_out/scale16k_$n.err || tail -5 gpurun_out/scale16k_$n.err
  fi
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29800 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/scale4k_8.json 2> gpurun_out/scale4k_8.err || tail -5 gpurun_out/scale4k_8.err
for f in scale16k_2 scale16k_4 scale16k_8 scale4k_8; do python -c "
import json
d=json.loads(open('gpurun_out/$f.json').read()); print('$f', d['n_gpus'], round(d['ms_per_step'],4), '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['roofline']['frac'], d['roofline'].get('nvlink'), d['clocks'])" 2>/dev/null; done
