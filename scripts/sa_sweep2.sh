#!/bin/bash
CW=${CW:-6:12 8:8 10:6 10:8 12:4 12:6 12:8 16:4 16:6}
for cw in $CW; do c=${cw%%:*}; w=${cw##*:};
  r=$(CHR3D_SA_CLUSTER=$c CHR3D_SA_WARPS=$w timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large-n 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['ms_per_step'], d['config']['cluster_size'], d['config']['threads_per_cta'])" 2>&1 | tail -1)
  echo "C=$c W=$w -> $r"
done
