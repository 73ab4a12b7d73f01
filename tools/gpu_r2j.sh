# grouped persistent z stage: parity + A/B timing (+ the ensemble e2e line if not yet measured)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_abi.py -m gpu -x -q -k "fused or persistent" 2>&1 | tail -5
echo "== grouped"; timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -4
echo "== classic"; MLV_ZADV_CLASSIC=1 timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3
echo "== grouped 2048"; timeout 300 python tools/kbench.py 2048 2048 50 2>&1 | head -3
echo "== classic 2048"; MLV_ZADV_CLASSIC=1 timeout 300 python tools/kbench.py 2048 2048 50 2>&1 | head -3
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 50 --warmup 5 --no-large-grid --no-cpu-baseline > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2j_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2j_bench.json'));print(d['ms_per_step'],d['value']);print(json.dumps(d['e2e'],indent=1)); print({k:v['ms'] for k,v in d['roofline']['kernels'].items()})"
