# multi-GPU session (run with gpurun --gpus N): NCCL / peer-memory parity of the sharded step + scaling bench
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "GPUs: $N"
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/pytest_sharded.log 2>&1; echo "pytest(p2p) rc=$?" >> gpurun_out/pytest_sharded.log
MLV_NO_P2P=1 timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q >> gpurun_out/pytest_sharded.log 2>&1; echo "pytest(nccl a2a) rc=$?" >> gpurun_out/pytest_sharded.log
grep -E "passed|failed|rc=|Error" gpurun_out/pytest_sharded.log | tail -8
for n in 2 4 8; do
  if [ $n -le $N ]; then
    for mode in p2p a2a; do
      if [ $mode = a2a ]; then export MLV_NO_P2P=1; else unset MLV_NO_P2P; fi
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29800 bench.py --gpus $n --steps 100 --warmup 10 > gpurun_out/scale_${n}_$mode.json 2> gpurun_out/scale_${n}_$mode.err
      echo "n=$n $mode rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/scale_${n}_$mode.json').read()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['nvlink'])" || tail -5 gpurun_out/scale_${n}_$mode.err
    done
  fi
done
