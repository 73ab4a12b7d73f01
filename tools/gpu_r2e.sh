# round 2, call e (2 GPUs): peer-store exchange ordered by device-side arrival counters
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "p2p and (tg64 or khlong or tearing)" > gpurun_out/r2e_pytest_sharded.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest_sharded.log; tail -4 gpurun_out/r2e_pytest_sharded.log
for b in flags nccl; do
MLV_P2P_BARRIER=$b timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 200 --warmup 10 --no-large-grid > gpurun_out/r2e_bench2_$b.json 2> gpurun_out/r2e_bench2_$b.err; echo "bench $b rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/r2e_bench2_$b.json'));print('$b',d['ms_per_step'],d['value'],d['parity']['ok'])"
done
