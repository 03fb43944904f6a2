#!/bin/bash
CFGS=${CFGS:-6:12:0 8:8:113000 8:6:113000 8:10:113000 10:6:113000 10:8:113000 12:6:113000 12:8:113000 6:8:113000}
for cfg in $CFGS; do IFS=: read c w cap <<< "$cfg"
  if [ "$cap" != "0" ]; then export CHR3D_SA_SMEM_CAP=$cap; else unset CHR3D_SA_SMEM_CAP; fi
  r=$(CHR3D_SA_CLUSTER=$c CHR3D_SA_WARPS=$w timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large-n 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['ms_per_step'], d['config']['cluster_size'], d['config']['threads_per_cta'])" 2>&1 | tail -1)
  echo "C=$c W=$w cap=$cap -> $r"
done
