# ncu evidence for configs 3 and 5: full captures of the FDM-z step kernels at 4096x2048 and of the
# long-line kernels on strips with 16384-point lines (full 16384^2 grids make ncu's save/restore too slow).
# The reports are summarised on the box (raw csv + markdown) and deleted: gpurun_out/ is capped at 64 MiB.
mkdir -p gpurun_out
cap() {  # cap <tag> <regex> <skip> <count> <command...>
    tag=$1; rx=$2; sk=$3; ct=$4; shift 4
    timeout 140 ncu --set full --clock-control none -k regex:"$rx" -s $sk -c $ct -f -o /tmp/prof_$tag "$@" > gpurun_out/r02d_ncu_$tag.log 2>&1
    echo "$tag rc=$?"
    ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv > gpurun_out/r02d_ncu_$tag.csv 2>/dev/null
    python tools/ncu_summary.py /tmp/prof_$tag.ncu-rep > gpurun_out/r02d_ncu_$tag.md 2>/dev/null
    rm -f /tmp/prof_$tag.ncu-rep
    wc -c gpurun_out/r02d_ncu_$tag.md
}
cap rbc "k_fdm|k_x1d|k_integrate" 25 5 python bench.py --config rbc --steps 4 --warmup 3 --no-cpu-baseline
cap x16k "k_xinv_split|k_xfwd" 12 4 python tools/bench_tearing.py --nx 16384 --nz 2048 --steps 3 --warmup 3
cap z16k "k_zr_advect" 8 2 python tools/bench_tearing.py --nx 2048 --nz 16384 --steps 3 --warmup 3
du -sh gpurun_out
