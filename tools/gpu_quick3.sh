# GPU session: parity tests + kernel micro-bench at 4096^2, 16384^2, 8192^2 + bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python tools/kbench.py > gpurun_out/kbench.log 2>&1; cat gpurun_out/kbench.log
timeout 300 python tools/kbench.py 16384 16384 5 > gpurun_out/kbench16k.log 2>&1; cat gpurun_out/kbench16k.log
timeout 300 python tools/kbench.py 8192 8192 10 > gpurun_out/kbench8k.log 2>&1; cat gpurun_out/kbench8k.log
timeout 600 python bench.py ${BENCH_ARGS:---steps 200 --warmup 10 --no-cpu-baseline} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
