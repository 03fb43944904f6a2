#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_persist.py tests/test_gpu_large_n.py tests/test_gpu_parity2.py -m gpu -q --timeout 120 > gpurun_out/pytest_persist.log 2>&1; echo "persist rc=$?"; tail -3 gpurun_out/pytest_persist.log
timeout 150 python scripts/ln_cta_fit.py > gpurun_out/r02_cta_fit.log 2>&1; echo "fit rc=$?"; cat gpurun_out/r02_cta_fit.log
timeout 200 python bench.py --workload synth_10kb_single --steps 2 --warmup 1 > gpurun_out/r02_bench_10kb.json 2> gpurun_out/r02_bench_10kb.err; echo "10kb rc=$?"
timeout 200 python bench.py --workload synth_50kb_single --steps 3 --warmup 2 > gpurun_out/r02_bench_50kb.json 2> gpurun_out/r02_bench_50kb.err; echo "50kb rc=$?"
python - <<'PY'
import json
for f in ["r02_bench_10kb.json","r02_bench_50kb.json"]:
    try:
        d=json.loads(open("gpurun_out/"+f).read().strip().split("\n")[-1])
        print(f, "ms/step", round(d["ms_per_step"],2), "evals", d["config"]["force_evals_per_model"], "eval_ms", round(d["force_eval_ms"],5), "tile_ms", round(d["tile_kernel_ms"],5), {k:(round(v,2) if isinstance(v,float) else v) for k,v in d["phases_ms_per_step"].items() if k!="note"}, "frac", round(d["roofline"]["frac"],4), round(d["roofline"]["frac_in_persistent_kernel"],4))
    except Exception as ex:
        print(f, "failed", ex)
PY
