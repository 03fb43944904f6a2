#!/bin/bash
# 8-GPU evidence, short form: ensemble weak scaling and the 24,925-bead structure with the fused exchange.
N=${N:-8}
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
run --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_scale_default_n$N.json 2> gpurun_out/r01c_scale_default_n$N.err; echo rc=$?
run --workload synth_10kb_single --steps 2 --warmup 2 > gpurun_out/r01c_scale_10kb_p2p_n$N.json 2> gpurun_out/r01c_scale_10kb_p2p_n$N.err; echo rc=$?
python - <<PY
import json
for f in ("default", "10kb_p2p"):
    try:
        d = json.loads(open("gpurun_out/r01c_scale_%s_n$N.json" % f).read().strip().split("\n")[-1])
        print(f, d["value"], d["ms_per_step"], d.get("force_eval_ms"), d.get("tile_kernel_ms"))
    except Exception as e:
        print(f, "ERR", e)
PY
