mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_abi.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
echo "TMA on"; timeout 120 python tools/kbench.py 4096 4096 20 2>&1 | head -3
echo "TMA off"; MLV_NO_TMA=1 timeout 120 python tools/kbench.py 4096 4096 20 2>&1 | head -3
