# N GPUs: tearing stepper with split joins (copy-engine exchange): parity (N=2) or timing (N=8)
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "tearing" 2>&1 | tail -2
fi
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 tools/bench_tearing.py --steps 10 --warmup 3 > gpurun_out/r3j_${N}_$tag.json 2> gpurun_out/r3j_${N}_$tag.err; echo "$tag N=$N rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r3j_${N}_$tag.json'));print(d['ms_per_step'])")"; }
run split MLV_DUMMY=1
run joined MLV_SPLIT_JOIN=0
