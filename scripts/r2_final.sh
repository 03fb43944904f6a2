#!/bin/bash
# Round-2 end-state visit of one B200: parity tests, bench lines, sparse provider timing, ncu evidence (launch list + one --set full
# capture per hot kernel, each after a plain run of the same command that exited 0).  Outputs: gpurun_out/${TAG}_*.
# LEAN=1: tests, default bench, reference arm and smoke only; NO_NCU=1: skip the profiler passes.
TAG=${TAG:-r02e}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out/${TAG}
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > ${O}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 ${O}_pytest_gpu.log
timeout 600 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > ${O}_bench_reference.json 2> ${O}_bench_reference.err; echo "reference rc=$?"
[ -z "$LEAN" ] && timeout 300 python bench.py --workload synth_10kb_single --steps 3 --warmup 1 > ${O}_bench_synth_10kb_single.json 2> ${O}_bench_10kb.err; echo "10kb rc=$?"
[ -z "$LEAN" ] && timeout 300 python bench.py --workload synth_50kb_single --steps 5 --warmup 2 > ${O}_bench_synth_50kb_single.json 2> ${O}_bench_50kb.err; echo "50kb rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > ${O}_smoke.log 2>&1; echo "smoke rc=$?"
[ -z "$LEAN" ] && timeout 600 python scripts/sparse_bench.py > ${O}_sparse_bench.jsonl 2> ${O}_sparse_bench.err; echo "sparse rc=$?"; cat ${O}_sparse_bench.jsonl
if [ -z "$NO_NCU" ]; then
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${O}_plain_bench.json 2> ${O}_plain_bench.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > ${O}_ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sa_kernel -s 3 -c 1 -o ${O}_sa_kernel -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-large-n > ${O}_ncu_sa.log 2>&1; echo "ncu sa rc=$?"
timeout 300 python scripts/large_n_probe.py 24925 > ${O}_plain_large_n.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_tile_kernel -s 26 -c 1 -o ${O}_ln_tile_kernel -f \
    python scripts/large_n_probe.py 24925 > ${O}_ncu_ln.log 2>&1; echo "ncu ln rc=$?"
timeout 300 python scripts/persist_probe.py > ${O}_plain_persist.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_persist_kernel -s 1 -c 1 -o ${O}_ln_persist_kernel -f \
    python scripts/persist_probe.py > ${O}_ncu_persist.log 2>&1; echo "ncu persist rc=$?"
fi
python - <<PY
import json
for f in ["bench_default", "bench_reference", "bench_synth_10kb_single", "bench_synth_50kb_single"]:
    try:
        d = json.loads([l for l in open("${O}_%s.json" % f).read().strip().split("\n") if l.startswith("{")][-1])
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "e2e", "roofline", "cpu_baseline", "gpu_launches", "clocks")})
        if "single_structure" in d: print("  single_structure", {k: d["single_structure"].get(k) for k in ("ms_per_step", "force_eval_ms", "tile_kernel_ms", "roofline", "error")})
    except Exception as ex:
        print(f, "failed", ex)
PY
ls -la gpurun_out/ | tail -30
