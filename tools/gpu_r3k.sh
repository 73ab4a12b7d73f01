mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tools/bench_tearing.py --steps 10 --warmup 3 > gpurun_out/r3k_split.json 2> gpurun_out/r3k_split.err; echo "split N=8 rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r3k_split.json'));print(d['ms_per_step'])")"
