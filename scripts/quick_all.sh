mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_exit=$?"
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python scripts/each_matrix.py 20 2>&1 | grep -c "   ok"
timeout 900 python scripts/each_matrix.py 20 2>&1 | grep -i fail
bash scripts/quick_batch.sh
