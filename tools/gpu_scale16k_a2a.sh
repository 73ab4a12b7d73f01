mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
export MLV_NO_P2P=1
for n in 2 8; do
  if [ $n -le $N ]; then
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29800 bench.py --gpus $n --steps 10 --warmup 3 --nx 16384 --nz 16384 > gpurun_out/scale16k_a2a_$n.json 2> gpurun_out/scale16k_a2a_$n.err || tail -5 gpurun_out/scale16k_a2a_$n.err
    python -c "
import json
d=json.loads(open('gpurun_out/scale16k_a2a_$n.json').read()); print('a2a', d['n_gpus'], round(d['ms_per_step'],4), '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'])"
  fi
done
