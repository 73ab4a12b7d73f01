# persistent inverse x pass with early request of the next source tile: parity + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_abi.py -m gpu -x -q -k "column_serial or fused or transforms_2d or full_size" 2>&1 | tail -3
echo "== persistent"; timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3
echo "== one tile per CTA"; MLV_XINV_ONESHOT=1 timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -1
echo "== 2048 persistent"; timeout 300 python tools/kbench.py 2048 2048 50 2>&1 | head -1
echo "== 2048 one tile per CTA"; MLV_XINV_ONESHOT=1 timeout 300 python tools/kbench.py 2048 2048 50 2>&1 | head -1
