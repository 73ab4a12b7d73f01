# A/B of the round-2 late variants (one GPU): fused 1-D advection with one packed line per CTA (config 3),
# split x passes with the shared-memory stash (16384-point lines, strip 16384 x 2048), parity of the variants.
# Needs tools/patches/r02d_split_stash_x1d_cols.patch applied (git apply) and the library rebuilt: the variants
# (MLV_X1D_COLS, MLV_XINV_STASH, MLV_XFWD_STASH) measured slower and were not merged (profiles/r02_experiments.md).
mkdir -p gpurun_out
rbc() { MLV_X1D_COLS=$1 timeout 120 python bench.py --config rbc --steps 50 --warmup 5 --no-cpu-baseline 2> gpurun_out/r4b_rbc_c$1.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('rbc cols=$1', d['ms_per_step'], d['roofline']['frac'])"; }
rbc 2
rbc 1
strip() { MLV_XINV_STASH=$1 MLV_XFWD_STASH=$2 timeout 120 python tools/bench_tearing.py --nx 16384 --nz 2048 --steps 6 --warmup 3 2> gpurun_out/r4b_strip_$1$2.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tearing 16384x2048 xinv_stash=$1 xfwd_stash=$2', d['ms_per_step'])"; }
strip 0 0
strip 0 1
strip 1 1
timeout 300 python -m pytest tests/test_gpu_abi.py -m gpu -x -q -k "split_lines_forced or columns_per_cta" 2>&1 | tail -2
MLV_XINV_STASH=1 MLV_XFWD_STASH=1 timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi.py -m gpu -x -q -k "16384 or long_line or tearing" 2>&1 | tail -2
