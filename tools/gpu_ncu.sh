# ncu full capture of the three step kernels (one launch each) on the bench workload
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_xinv|k_z_advect|k_xfwd" -s 9 -c 3 -f -o gpurun_out/prof python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
