echo "== default (C=2, 1 CTA/SM for x passes)"
python tools/kbench.py 2>&1 | grep -E "x_inv|x_forw|advect|z_"
echo "== XC1 variant (C=1, 2 CTAs/SM for x passes)"
cp melvin.py_b200/melvin/_lib/libmelvin_b200.so /tmp/keep.so
cp melvin.py_b200/melvin/_lib/libmelvin_b200_xc1.so melvin.py_b200/melvin/_lib/libmelvin_b200.so
python tools/kbench.py 2>&1 | grep -E "x_inv|x_forw|advect|z_"
python -m pytest tests/test_gpu_abi.py -m gpu -x -q -k "fused or transforms_2d" 2>&1 | tail -2
cp /tmp/keep.so melvin.py_b200/melvin/_lib/libmelvin_b200.so
