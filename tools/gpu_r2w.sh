# specialised kernels: z stage without sharding code, single-scalar forward x pass (UN sweep)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_abi.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
echo "== default"; timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3
echo "== z generic"; MLV_ZADV_GENERIC=1 timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -2 | tail -1
echo "== xfwd generic"; MLV_XFWD_GENERIC=1 timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3 | tail -1
for u in 2 4 6; do echo "== xfwd UN $u"; MLV_XFWD_UN=$u timeout 300 python tools/kbench.py 4096 4096 50 2>&1 | head -3 | tail -1; done
