echo "== co-resident CTAs allowed (CHR3D_SA_MIN_SMEM=0), concurrent groups, x4 processes"
for i in 1 2 3 4; do CHR3D_SA_MIN_SMEM=0 timeout 300 python scripts/all_probe.py 20 | tail -4 | tr '\n' ' '; echo; done
echo "== 12x4 forced, co-resident, x4"
for i in 1 2 3 4; do CHR3D_SA_MIN_SMEM=0 CHR3D_SA_CLUSTER=12 CHR3D_SA_WARPS=4 timeout 120 python scripts/one_matrix.py chr4_500kb 20 40 | tail -1; done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-large-n | python -c "import sys,json; r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default', r['ms_per_step'], r['value'], r['e2e']['value'])"
