#!/bin/bash
# Second short GPU visit (2.4 GPU-minutes left): the tests that touch the changed code (parallel sums with the loads issued up front,
# one-pass list building, warp scan), the default bench line, the same-box A/B and the launch list of one call, the host trace.
TAG=${TAG:-r02j}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/${TAG}
date +%s > ${O}_t0
timeout 60 python -m pytest tests/test_gpu_seqsum.py tests/test_gpu_parity.py tests/test_perl_frontend.py -m gpu -q --timeout 60 \
    -k "seq_sum or mean_of or restraint or assess or coil or many_jobs or reproducible or front_end_whole or benchmarked" > ${O}_pytest_subset.log 2>&1; echo "pytest rc=$?"; tail -4 ${O}_pytest_subset.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${O}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 ${O}_smoke.log
timeout 60 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err; echo "bench rc=$?"
timeout 30 python scripts/e2e_once.py 5 > ${O}_e2e_ab.json 2> ${O}_e2e_ab.err; echo "e2e A/B rc=$?"; cat ${O}_e2e_ab.json
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_e2e.csv \
    python scripts/e2e_once.py 1 > ${O}_ncu_e2e.log 2>&1; echo "launch list rc=$?"
CHR3D_LIBRARY=$PWD/chromosome3d_b200/libchr3d_dev.so CHR3D_TRACE=1 timeout 20 python scripts/e2e_trace.py > ${O}_trace.log 2>&1; echo "trace rc=$?"; tail -12 ${O}_trace.log
python - <<PY
import json
try:
    d = json.loads([l for l in open("${O}_bench_default.json").read().strip().split("\n") if l.startswith("{")][-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "e2e", "gpu_launches")})
except Exception as ex:
    print("failed", ex)
PY
echo "elapsed $(( $(date +%s) - $(cat ${O}_t0) )) s"
