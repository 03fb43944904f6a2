#!/bin/bash
# A/B of the barrier-free middle MD evaluations (dev build, CHR3D_LN_FLOW) on N GPUs: bash scripts/r2_ab_flow.sh <ngpus>
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for flow in 1 0 1; do
CHR3D_LIBRARY=$PWD/chromosome3d_b200/libchr3d_dev.so CHR3D_LN_FLOW=$flow timeout 200 $TR --master-port 2952$flow bench.py --gpus $N --workload synth_10kb_single --steps 2 --warmup 1 2> gpurun_out/ab_flow.err | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().split('\n')[-1])
print('flow=$flow n=$N ms/step %.1f eval_ms %.4f tile_kernel_ms %.4f' % (d['ms_per_step'], d['force_eval_ms'], d['tile_kernel_ms']), {k: round(v, 1) for k, v in d['phases_ms_per_step'].items() if isinstance(v, (int, float))})" | tee -a gpurun_out/r02f_ab_flow_n$N.log
done
