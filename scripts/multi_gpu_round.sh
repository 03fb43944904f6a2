#!/bin/bash
# N-GPU visit: multi-GPU test, ensemble weak scaling and single-structure strong scaling bench lines (P2P and NCCL exchange).
N=${N:-2}
nvidia-smi -L | head -2
if [ "$N" = "2" ]; then timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -8; fi
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
run --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_default_n$N.json 2> gpurun_out/scale_default_n$N.err; echo rc=$?; cut -c1-300 gpurun_out/scale_default_n$N.json
run --workload synth_10kb_single --steps 2 --warmup 2 > gpurun_out/scale_10kb_p2p_n$N.json 2> gpurun_out/scale_10kb_p2p_n$N.err; echo rc=$?; python -c "
import json;d=json.loads(open('gpurun_out/scale_10kb_p2p_n$N.json').read().strip().split('\n')[-1]);print('p2p', d['value'],d['ms_per_step'],d['force_eval_ms'],d['tile_kernel_ms'])"; tail -2 gpurun_out/scale_10kb_p2p_n$N.err
CHR3D_NO_P2P=1 run --workload synth_10kb_single --steps 2 --warmup 2 > gpurun_out/scale_10kb_nccl_n$N.json 2> gpurun_out/scale_10kb_nccl_n$N.err; echo rc=$?; python -c "
import json;d=json.loads(open('gpurun_out/scale_10kb_nccl_n$N.json').read().strip().split('\n')[-1]);print('nccl', d['value'],d['ms_per_step'],d['force_eval_ms'],d['tile_kernel_ms'])"
run --workload synth_50kb_single --steps 2 --warmup 2 > gpurun_out/scale_50kb_p2p_n$N.json 2> gpurun_out/scale_50kb_p2p_n$N.err; echo rc=$?; python -c "
import json;d=json.loads(open('gpurun_out/scale_50kb_p2p_n$N.json').read().strip().split('\n')[-1]);print('p2p 50kb', d['value'],d['ms_per_step'],d['force_eval_ms'],d['tile_kernel_ms'])"
