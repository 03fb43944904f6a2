#!/bin/bash
# Development: sweep the SA kernel launch shape on the bench workload (device-timed ms per step).
for c in ${CS:-4 5 6}; do for w in ${WS:-4 8 12 16}; do
  r=$(CHR3D_SA_CLUSTER=$c CHR3D_SA_WARPS=$w timeout 120 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['ms_per_step'], d['config']['cluster_size'], d['config']['threads_per_cta'], d['config']['force_evals_per_model'])" 2>&1 | tail -1)
  echo "C=$c W=$w -> $r"
done; done
