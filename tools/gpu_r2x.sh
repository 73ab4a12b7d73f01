# specialised forward x pass: resident CTAs with early request of the next staged block
mkdir -p gpurun_out
MLV_XFWD_PERSISTENT=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "baseline_size_4096 or single_scalar or config1" 2>&1 | tail -2
echo "== one tile per CTA"; timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3 | tail -1
echo "== resident"; MLV_XFWD_PERSISTENT=1 timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3 | tail -1
echo "== 2048 one"; timeout 300 python tools/kbench.py 2048 2048 50 2>&1 | head -3 | tail -1
echo "== 2048 resident"; MLV_XFWD_PERSISTENT=1 timeout 300 python tools/kbench.py 2048 2048 50 2>&1 | head -3 | tail -1
