#!/bin/bash
# the whole GPU parity suite on one B200
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/r02h_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02h_pytest_gpu.log
