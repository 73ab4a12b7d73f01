# 8 GPUs: tearing 16384^2, knobs of the copy-engine exchange
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711 tools/bench_tearing.py --steps 10 --warmup 3 > gpurun_out/r2z_$tag.json 2> gpurun_out/r2z_$tag.err; echo "$tag rc=$? $(cut -c1-300 gpurun_out/r2z_$tag.json)"; }
run default MLV_DUMMY=1
run streams8 MLV_COPY_STREAMS=8
run chunks2 MLV_FWD_CHUNKS=2
run chunks8 MLV_FWD_CHUNKS=8 MLV_COPY_STREAMS=8
