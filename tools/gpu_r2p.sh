# forward x pass: prefetch of the next CTA's operand blocks + tensor prefetch of the epilogue columns
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_abi.py -m gpu -x -q -k "column_serial or fused or transforms_2d or full_size" 2>&1 | tail -3
echo "== tensor prefetch"; timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3
echo "== per-row hints"; MLV_XFWD_LINEPF=1 timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3 | tail -1
echo "== 2048"; timeout 300 python tools/kbench.py 2048 2048 50 2>&1 | head -3
timeout 900 python bench.py --steps 50 --warmup 5 --no-large-grid --no-cpu-baseline > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2p_bench.err
python -c "
import json;d=json.load(open('gpurun_out/r2p_bench.json'));print(d['ms_per_step'],d['value'],d['roofline']['step']['frac']); print({k:v['ms'] for k,v in d['roofline']['kernels'].items()}); print(d['e2e']['ms_per_step'], d['e2e']['value'])"
