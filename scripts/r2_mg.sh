#!/bin/bash
# round-2 multi-GPU visit: bash scripts/r2_mg.sh <ngpus>
# trajectory equality of the sharded single structure against one GPU, the 24,925-bead structure over the ranks, the default line
N=${1:-2}
QUICK=${2:-}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_mg${N}_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
CHR3D_TEST_P2P=1 timeout 200 $TR --master-port 29511 scripts/mg_worker.py > gpurun_out/r02_mg${N}_equality.log 2>&1; echo "equality rc=$?"; grep RESULT gpurun_out/r02_mg${N}_equality.log | cut -c1-1200
timeout 200 $TR --master-port 29512 bench.py --gpus $N --workload synth_10kb_single --steps 2 --warmup 1 > gpurun_out/r02_scale_10kb_n${N}.json 2> gpurun_out/r02_scale_10kb_n${N}.err; echo "10kb rc=$?"
timeout 200 $TR --master-port 29513 bench.py --gpus $N --workload synth_50kb_single --steps 3 --warmup 2 > gpurun_out/r02_scale_50kb_n${N}.json 2> gpurun_out/r02_scale_50kb_n${N}.err; echo "50kb rc=$?"
if [ "$N" = "2" ] && [ -z "$QUICK" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 800 > gpurun_out/r02_mg2_pytest.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/r02_mg2_pytest.log
  [ -n "$WITH_NCCL" ] && { CHR3D_NO_P2P=1 timeout 900 $TR --master-port 29514 bench.py --gpus $N --workload synth_10kb_single --steps 1 --warmup 1 > gpurun_out/r02_scale_10kb_nccl_n${N}.json 2> gpurun_out/r02_scale_10kb_nccl_n${N}.err; echo "10kb nccl rc=$?"; }
fi
[ -z "$QUICK" ] && timeout 1200 $TR --master-port 29515 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_scale_default_n${N}.json 2> gpurun_out/r02_scale_default_n${N}.err; echo "default rc=$?"
python - <<PY
import json
for f in ["r02_scale_10kb_n$N.json","r02_scale_50kb_n$N.json","r02_scale_10kb_nccl_n$N.json"]:
    try:
        d=json.loads([l for l in open("gpurun_out/"+f).read().strip().split("\n") if l.startswith("{")][-1])
        print(f, "ms/step", round(d["ms_per_step"],2), "evals", d["config"]["force_evals_per_model"], "eval_ms", round(d["force_eval_ms"],5), "tile_ms", round(d["tile_kernel_ms"],5), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["phases_ms_per_step"].items() if k!="note"}, "frac", round(d["roofline"]["frac"],4))
    except Exception as ex:
        print(f, "failed", ex)
try:
    d=json.loads([l for l in open("gpurun_out/r02_scale_default_n$N.json").read().strip().split("\n") if l.startswith("{")][-1])
    ss=d.get("single_structure",{})
    print("default: value", d["value"], "ms", d["ms_per_step"], "| single_structure ms", ss.get("ms_per_step"), ss.get("error"))
except Exception as ex:
    print("default failed", ex)
PY
tail -c 600 gpurun_out/r02_scale_10kb_n${N}.err
