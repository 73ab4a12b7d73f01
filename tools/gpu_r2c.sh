# round 2, call c (2 GPUs): sharded parity tests + the driver's multi-GPU bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r2c_pytest_sharded.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_sharded.log; tail -4 gpurun_out/r2c_pytest_sharded.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/r2c_bench2.json 2> gpurun_out/r2c_bench2.err ) 2>&1 | tail -3; echo "bench rc=$?"; tail -4 gpurun_out/r2c_bench2.err; cut -c1-300 gpurun_out/r2c_bench2.json
