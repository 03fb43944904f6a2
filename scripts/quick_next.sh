mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_next_rows.py tests/test_perl_frontend.py -m gpu -x -q > gpurun_out/pytest_next.log 2>&1; echo "pytest_exit=$?"
tail -25 gpurun_out/pytest_next.log
