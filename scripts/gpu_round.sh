#!/bin/bash
# One GPU-box visit: parity tests, bench, large-N probe.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_exit=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python scripts/large_n_probe.py 4985 24925 > gpurun_out/large_n_probe.log 2>&1; echo "probe_exit=$?"
cat gpurun_out/large_n_probe.log | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench_exit=$?"
tail -c 3000 gpurun_out/bench.log
