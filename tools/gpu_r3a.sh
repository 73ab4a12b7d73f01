# 2 GPUs: sharded parity after the kernel changes; copies by CTAs (MLV_COPY_CTAS) vs copy engines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q > gpurun_out/r3a_sharded_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3a_sharded_tests.log
MLV_COPY_CTAS=8 timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "dma" > gpurun_out/r3a_sharded_tests_ctas.log 2>&1; echo "pytest(ctas) rc=$?"; tail -3 gpurun_out/r3a_sharded_tests_ctas.log
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/bench_tearing.py --steps 6 --warmup 3 > gpurun_out/r3a_$tag.json 2> gpurun_out/r3a_$tag.err; echo "$tag rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r3a_$tag.json'));print(d['ms_per_step'])")"; }
run engines MLV_DUMMY=1
run ctas8 MLV_COPY_CTAS=8
run ctas16 MLV_COPY_CTAS=16
