# members of the streamed e2e ensemble: 2 / 6 / 8 against the default 4
for m in 2 6 8 4; do
MLV_E2E_MEMBERS=$m timeout 30 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-large-grid 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('members $m', d['e2e']['ms_per_step'], d['e2e']['blocking']['ms_per_step'], d['ms_per_step'])"
done
