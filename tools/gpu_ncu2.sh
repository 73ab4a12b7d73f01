mkdir -p gpurun_out
# launch list (per-launch durations, serialised) and one full capture of each step kernel
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_xinv|k_z_advect|k_xfwd" -s 9 -c 3 -o gpurun_out/prof_r01c -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/prof_r01c.ncu-rep
