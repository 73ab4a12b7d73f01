# 8-GPU session: sharded parity (peer memory) + strong-scaling bench 1,2,4,8
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/scale8_1.json 2> gpurun_out/scale8_1.err
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29800 bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/scale8_$n.json 2> gpurun_out/scale8_$n.err || tail -5 gpurun_out/scale8_$n.err
  fi
done
for n in 1 2 4 8; do python -c "
import json
d=json.loads(open('gpurun_out/scale8_$n.json').read()); print(d['n_gpus'], round(d['ms_per_step'],4), '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], d['clocks'])" 2>/dev/null; done
