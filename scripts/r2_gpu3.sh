#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_persist.py -m gpu -q --timeout 600 > gpurun_out/pytest_persist.log 2>&1; echo "persist rc=$?"; tail -3 gpurun_out/pytest_persist.log
bash scripts/r2_mg.sh 2 quick
