echo "== hints"; timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3 | tail -1
echo "== tensor prefetch"; MLV_XFWD_TENSORPF=1 timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3 | tail -1
echo "== no prefetch of the columns"; MLV_XFWD_TENSORPF=1 MLV_NO_TMA_X=1 timeout 300 python tools/kbench.py 2048 2048 50 2>&1 | head -3 | tail -1
MLV_XFWD_TENSORPF=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "baseline_size_4096 or single_scalar" 2>&1 | tail -1
