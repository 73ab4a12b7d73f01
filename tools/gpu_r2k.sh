# A/B of z-stage variants at 4096^2 (kbench) + quick parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_abi.py -m gpu -x -q -k "fused or persistent or transforms_2d" 2>&1 | tail -3
echo "== grouped"; timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -5
echo "== classic"; MLV_ZADV_CLASSIC=1 timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3
