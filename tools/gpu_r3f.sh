timeout 900 python -m pytest tests/test_gpu_abi.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
echo "== 8192 tensor pf"; timeout 300 python tools/kbench.py 8192 8192 20 2>&1 | head -3 | tail -1
echo "== 8192 hints"; MLV_XFWD_LINEPF=1 timeout 300 python tools/kbench.py 8192 8192 20 2>&1 | head -3 | tail -1
echo "== 16384 tensor pf"; timeout 300 python tools/kbench.py 16384 16384 10 2>&1 | head -3 | tail -1
echo "== 16384 hints"; MLV_XFWD_LINEPF=1 timeout 300 python tools/kbench.py 16384 16384 10 2>&1 | head -3 | tail -1
