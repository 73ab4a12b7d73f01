# round 2, call f: x passes with independent lines (named barriers) vs lane-interleaved columns
mkdir -p gpurun_out
for m in 0 1 2 3; do echo "MLV_XBLK=$m"; MLV_XBLK=$m python tools/kbench.py 4096 4096 50 2>&1 | grep -v "torch copy\|z_inverse\|z_forward"; done | tee gpurun_out/r2f_xblk.txt
MLV_XBLK=3 timeout 600 python -m pytest tests/test_gpu_abi.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
