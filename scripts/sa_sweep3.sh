#!/bin/bash
# Development: launch-shape sweep on the throughput workload (400 trajectories of chr1_500kb)
WL=${WL:-chr1_500kb_alpha_sweep_m20}
CFGS=${CFGS:-2:16:0 2:20:0 3:12:0 3:16:0 4:8:110000 4:12:110000 3:8:110000 4:16:0 6:8:110000}
for cfg in $CFGS; do IFS=: read c w cap <<< "$cfg"
  if [ "$cap" != "0" ]; then export CHR3D_SA_SMEM_CAP=$cap; else unset CHR3D_SA_SMEM_CAP; fi
  r=$(CHR3D_SA_CLUSTER=$c CHR3D_SA_WARPS=$w timeout 200 python bench.py --workload $WL --steps 1 --warmup 1 --no-cpu-baseline --no-large-n 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['ms_per_step'], d['value'], d['config']['cluster_size'], d['config']['threads_per_cta'])" 2>&1 | tail -1)
  echo "C=$c W=$w cap=$cap -> $r"
done
