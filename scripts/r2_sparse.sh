#!/bin/bash
# sparse provider parity + quick regression of the persistent-kernel tests
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_sparse.py -x -q > gpurun_out/pytest_sparse.log 2>&1; echo "sparse rc=$?"; tail -25 gpurun_out/pytest_sparse.log
