#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/lbfgs_track.py > gpurun_out/lbfgs_track2.log 2>&1; echo "track rc=$?"; grep "drop 1e+01" gpurun_out/lbfgs_track2.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_parity.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/pytest_parity.log
