# column-serial inverse x pass: parity + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_abi.py -m gpu -x -q -k "column_serial or fused or transforms_2d or full_size" 2>&1 | tail -3
echo "== cols"; timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3
echo "== classic xinv"; MLV_XINV_CLASSIC=1 timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3
echo "== cols grid 273"; MLV_XINV_GRID=273 timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -1
echo "== cols grid 148"; MLV_XINV_GRID=148 timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -1
