#!/bin/bash
# Last GPU visit of round 2 (4.8 GPU-minutes left): the whole parity suite with the parallel left-to-right sums and the staged coil
# kernel in, the default bench line, the same-box A/B of the e2e call, the launch list of one call, smoke, and — if time is left —
# the 900-job sweep.  Every step under its own timeout; the essential ones first.
TAG=${TAG:-r02i}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/${TAG}
date +%s > ${O}_t0
timeout 160 python -m pytest tests -m gpu -q --timeout 120 > ${O}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 ${O}_pytest_gpu.log
timeout 90 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err; echo "bench rc=$?"
timeout 40 python scripts/e2e_once.py 5 > ${O}_e2e_ab.json 2> ${O}_e2e_ab.err; echo "e2e A/B rc=$?"; cat ${O}_e2e_ab.json
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_e2e.csv \
    python scripts/e2e_once.py 1 > ${O}_ncu_e2e.log 2>&1; echo "launch list rc=$?"
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${O}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 70 python bench.py --workload all_inputs_alpha_sweep_m20 --steps 1 --warmup 1 --no-cpu-baseline --no-large-n > ${O}_bench_sweep.json 2> ${O}_bench_sweep.err; echo "sweep rc=$?"
python - <<PY
import json
for f in ["bench_default", "bench_sweep"]:
    try:
        d = json.loads([l for l in open("${O}_%s.json" % f).read().strip().split("\n") if l.startswith("{")][-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "e2e", "gpu_launches")})
    except Exception as ex:
        print(f, "failed", ex)
PY
echo "elapsed $(( $(date +%s) - $(cat ${O}_t0) )) s"
