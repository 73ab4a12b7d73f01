echo "== ping-pong z stage"; python tools/kbench.py 2>&1 | grep -E "advect"
echo "== old z stage"; MLV_NO_PINGPONG=1 python tools/kbench.py 2>&1 | grep -E "advect"
