# round 2, session 2 first call: GPU parity tests, smoke, bench lines (KH 4096^2 + large_grid, reference arm, rbc), launch list + full ncu capture
TAG=${1:-r02a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv | tail -1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -5 gpurun_out/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ) 2>&1 | tail -3; echo "bench rc=$?"; tail -5 gpurun_out/${TAG}_bench.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err ) 2>&1 | tail -3; cut -c1-400 gpurun_out/${TAG}_bench_ref.json
timeout 600 python bench.py --config rbc --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_rbc.json 2> gpurun_out/${TAG}_bench_rbc.err; echo "rbc rc=$?"; cut -c1-600 gpurun_out/${TAG}_bench_rbc.json; tail -3 gpurun_out/${TAG}_bench_rbc.err
bash tools/gpu_r2d.sh ${TAG}
