#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_next_rows.py -x -q -m gpu > gpurun_out/pytest_next.log 2>&1; echo "next rc=$?"; tail -15 gpurun_out/pytest_next.log
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -s -k "shipped" > gpurun_out/pytest_pin.log 2>&1; echo "pin rc=$?"; grep -E "pin=|passed|failed|Error" gpurun_out/pytest_pin.log | tail
timeout 200 python scripts/lbfgs_track.py > gpurun_out/lbfgs_track.log 2>&1; echo "track rc=$?"; cat gpurun_out/lbfgs_track.log | tail -60
