#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() {   # tag, env...
  tag=$1; shift
  env "$@" timeout 200 $TR --master-port 29512 bench.py --gpus 8 --workload synth_10kb_single --steps 2 --warmup 1 > gpurun_out/r02_scale_10kb_n8_$tag.json 2> gpurun_out/r02_scale_10kb_n8_$tag.err; echo "$tag rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02_scale_10kb_n8_$tag.json").read().strip().split("\n") if l.startswith("{")][-1])
print("$tag: ms/step", round(d["ms_per_step"],2), "evals", d["config"]["force_evals_per_model"], "eval_us", round(1e3*d["force_eval_ms"],2), "tile_us", round(1e3*d["tile_kernel_ms"],2))
ph=d["phases_ms_per_step"]; ev=d["config"]["force_evals_per_model"]
print(ph["by_rank_columns"][3:])
for r in ph["by_rank"]:
    print([round(1e3*v/ev,1) for v in r[3:]])
PY
}
DEV=CHR3D_LIBRARY=$PWD/chromosome3d_b200/libchr3d_dev.so
run base X=1
run tail1 $DEV CHR3D_LN_RING_TAIL=1
run persist $DEV CHR3D_LN_L2_PERSIST=1
run noflow $DEV CHR3D_LN_FLOW=0
