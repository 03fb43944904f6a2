#!/bin/bash
# Third, shortest GPU visit (1.2 GPU-minutes left): the sums with the three-addition contribution — the tests that touch them, the
# same-box A/B of the e2e call (which also asserts identical results), smoke, and the default bench line if the time allows.
TAG=${TAG:-r02k}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/${TAG}
timeout 30 python -m pytest tests/test_gpu_seqsum.py tests/test_gpu_parity.py -m gpu -q --timeout 30 -x \
    -k "seq_sum or mean_of or restraint_table or assess or many_jobs" > ${O}_pytest_subset.log 2>&1; echo "pytest rc=$?"; tail -3 ${O}_pytest_subset.log
timeout 15 python scripts/e2e_once.py 3 > ${O}_e2e_ab.json 2> ${O}_e2e_ab.err; echo "e2e A/B rc=$?"; cat ${O}_e2e_ab.json
timeout 20 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${O}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 ${O}_smoke.log
timeout 40 python bench.py --no-large-n --no-cpu-baseline > ${O}_bench_default_nolarge.json 2> ${O}_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads([l for l in open("${O}_bench_default_nolarge.json").read().strip().split("\n") if l.startswith("{")][-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "e2e")})
except Exception as ex:
    print("failed", ex)
PY
