#!/bin/bash
# sparse provider: parity tests + per-evaluation timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_parity.py -x -q > gpurun_out/pytest_sparse.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/pytest_sparse.log
timeout 500 python scripts/sparse_bench.py > gpurun_out/r02f_sparse_bench.jsonl 2> gpurun_out/r02f_sparse_bench.err; echo "sparse rc=$?"; cat gpurun_out/r02f_sparse_bench.jsonl; tail -3 gpurun_out/r02f_sparse_bench.err
for k in 1; do timeout 300 python bench.py --no-cpu-baseline --no-large-n > gpurun_out/r02f_bench_default_$k.json 2> gpurun_out/r02f_bench_default_$k.err; echo "bench $k rc=$?"; python -c "
import json; d=json.loads([l for l in open('gpurun_out/r02f_bench_default_$k.json') if l.startswith('{')][-1]); print(d['ms_per_step'], d['e2e'])"; done
