python tools/kbench.py 2>&1 | grep -E "x_inv|x_forw|advect|z_"
