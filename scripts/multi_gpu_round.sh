#!/bin/bash
# 2-GPU visit: multi-GPU test, ensemble weak scaling and single-structure strong scaling bench lines.
N=${N:-2}
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@"; }
run --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_default_n$N.json 2> gpurun_out/scale_default_n$N.err; echo rc=$?; cut -c1-600 gpurun_out/scale_default_n$N.json; tail -3 gpurun_out/scale_default_n$N.err
run --workload synth_10kb_single --steps 2 --warmup 2 > gpurun_out/scale_10kb_n$N.json 2> gpurun_out/scale_10kb_n$N.err; echo rc=$?; cut -c1-900 gpurun_out/scale_10kb_n$N.json; tail -3 gpurun_out/scale_10kb_n$N.err
run --workload synth_50kb_single --steps 2 --warmup 2 > gpurun_out/scale_50kb_n$N.json 2> gpurun_out/scale_50kb_n$N.err; echo rc=$?; cut -c1-600 gpurun_out/scale_50kb_n$N.json
