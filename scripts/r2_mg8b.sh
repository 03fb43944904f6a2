#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29512 bench.py --gpus 8 --workload synth_10kb_single --steps 2 --warmup 1 > gpurun_out/r02_scale_10kb_n8.json 2> gpurun_out/r02_scale_10kb_n8.err; echo "10kb rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/r02_scale_10kb_n8.json").read().strip().split("\n") if l.startswith("{")][-1])
print("ms/step", d["ms_per_step"], "evals", d["config"]["force_evals_per_model"], "eval_ms", d["force_eval_ms"], "tile_ms", d["tile_kernel_ms"])
ph=d["phases_ms_per_step"]; ev=d["config"]["force_evals_per_model"]
print(ph["by_rank_columns"])
for r in ph["by_rank"]:
    print([round(v,3) if k<3 else round(1e3*v/ev,2) for k,v in enumerate(r)])
PY
