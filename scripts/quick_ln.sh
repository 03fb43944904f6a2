mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_large_n.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_exit=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python scripts/large_n_probe.py 4985 24925 > gpurun_out/large_n_probe.log 2>&1; echo "probe_exit=$?"
cat gpurun_out/large_n_probe.log | tail -8
