# A/B: exchange register order, static lo/hi split; ensemble depth
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_abi.py -m gpu -x -q -k "fused or persistent or transforms_2d" 2>&1 | tail -3
echo "== default grouped"; timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -5
echo "== default classic"; MLV_ZADV_CLASSIC=1 timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3
echo "== noxorder grouped"; MLV_LIB=$PWD/melvin.py_b200/melvin/_lib/variant_noxorder.so timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -5
echo "== noxorder classic"; MLV_ZADV_CLASSIC=1 MLV_LIB=$PWD/melvin.py_b200/melvin/_lib/variant_noxorder.so timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3
for m in 3 4 6; do
MLV_E2E_MEMBERS=$m timeout 600 python bench.py --steps 30 --warmup 5 --no-large-grid --no-cpu-baseline > gpurun_out/r2l_bench_m$m.json 2> gpurun_out/r2l_bench_m$m.err; echo "members $m rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/r2l_bench_m$m.json'));print(d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['value'], d['e2e']['blocking']['ms_per_step'])"
done
