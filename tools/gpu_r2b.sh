# round 2, call b: fused Fourier-x / FDM-z step (config 3) -- parity at size, bench, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_gpu.log; tail -5 gpurun_out/r2b_pytest_gpu.log
timeout 600 python bench.py --config rbc --steps 50 --warmup 5 > gpurun_out/r2b_bench_rbc.json 2> gpurun_out/r2b_bench_rbc.err; echo "rbc rc=$?"; cut -c1-900 gpurun_out/r2b_bench_rbc.json; tail -3 gpurun_out/r2b_bench_rbc.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 80 --csv --log-file gpurun_out/r2b_rbc_launches.csv python bench.py --config rbc --steps 5 --warmup 3 > /dev/null 2> gpurun_out/r2b_ncu.err; echo "ncu rc=$?"
