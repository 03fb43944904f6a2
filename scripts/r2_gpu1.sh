#!/bin/bash
# round-2 GPU visit, one GPU: persistent-kernel tests first (own process: a trap poisons the context), then the rest, then bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_gpu_persist.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_persist.log 2>&1; echo "persist rc=$?" | tee -a gpurun_out/pytest_persist.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_persist.py --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "gpu rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_persist.log gpurun_out/pytest_gpu.log
timeout 600 python bench.py --workload synth_10kb_single --steps 2 --warmup 1 > gpurun_out/r02_bench_10kb.json 2> gpurun_out/r02_bench_10kb.err; echo "10kb rc=$?"
timeout 600 python bench.py --workload synth_50kb_single --steps 3 --warmup 2 > gpurun_out/r02_bench_50kb.json 2> gpurun_out/r02_bench_50kb.err; echo "50kb rc=$?"
timeout 900 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "default rc=$?"
tail -c 1500 gpurun_out/r02_bench_10kb.json; tail -c 600 gpurun_out/r02_bench_10kb.err
