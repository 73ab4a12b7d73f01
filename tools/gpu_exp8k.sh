mkdir -p gpurun_out
for f in 0 1 2 3; do echo "MLV_FORCE_SPLIT=$f"; MLV_FORCE_SPLIT=$f timeout 300 python tools/kbench.py 8192 8192 10 2>&1 | head -5; done > gpurun_out/kbench8k_split.log 2>&1
cat gpurun_out/kbench8k_split.log
