# streamed-ensemble e2e: new GPU test + bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ensemble or async" 2>&1 | tail -5
timeout 900 python bench.py --steps 50 --warmup 5 --no-large-grid --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/r2i_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2i_bench.json'));print(d['ms_per_step'],d['value']);print(json.dumps(d['e2e'],indent=1))"
