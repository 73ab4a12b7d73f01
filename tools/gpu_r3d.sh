echo "== api order"; timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3
echo "== old order"; KBENCH_XINV_ORDER=old timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -1
