#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_persist.py tests/test_gpu_large_n.py -m gpu -q --timeout 120 > gpurun_out/pytest_persist.log 2>&1; echo "persist rc=$?"; tail -3 gpurun_out/pytest_persist.log
timeout 120 python scripts/ab_probe.py 2>&1 | tail -1
