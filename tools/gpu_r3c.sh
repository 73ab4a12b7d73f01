# 8 GPUs: tearing 16384^2 with the exchange done by CTAs
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tools/bench_tearing.py --steps 10 --warmup 3 > gpurun_out/r3c_$tag.json 2> gpurun_out/r3c_$tag.err; echo "$tag rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r3c_$tag.json'));print(d['ms_per_step'])")"; }
run ctas16 MLV_COPY_CTAS=16 MLV_COPY_STREAMS=8
run ctas8 MLV_COPY_CTAS=8 MLV_COPY_STREAMS=8
