mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_exit=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-large-n > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench_exit=$?"
python - <<'PY'
import json
r = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print({k: r[k] for k in ("value", "ms_per_step", "sa_models_per_sec")}, r["e2e"]["value"], r["roofline"]["frac"], r["config"]["cluster_size"], r["config"]["threads_per_cta"])
PY
