# round 2: launch list of the bench command + one full capture of each step kernel (KH 4096^2)
# usage: bash tools/gpu_r2d.sh <tag>
TAG=${1:-r02a}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_xinv|k_z_advect|k_xfwd|k_reduce" -s 62 -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-large-grid > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_xinv|k_z_advect|k_xfwd" -s 74 -c 3 -o gpurun_out/prof_${TAG} -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-large-grid > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/prof_${TAG}.ncu-rep
