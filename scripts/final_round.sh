#!/bin/bash
# End-of-round evidence: smoke, default bench (+ CPU baseline), reference arm, other workloads, ncu launch list + full captures.
TAG=${TAG:-r01c}
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke_exit=$?"; tail -1 gpurun_out/${TAG}_smoke.log | cut -c1-200
python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; echo "bench_exit=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref_exit=$?"
for w in chr1_500kb_alpha_sweep_m20 all_inputs_a1.1_m20 synth_10kb_single synth_50kb_single; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$w.json 2> gpurun_out/${TAG}_bench_$w.err; echo "$w exit=$?"
done
TAG=$TAG bash scripts/ncu_round.sh > gpurun_out/${TAG}_ncu_round.log 2>&1; echo "ncu_exit=$?"
for d in 1 2; do CHR3D_LN_DEBUG=$d python scripts/large_n_probe.py 24925 2>&1 | grep -v fp32_peak > gpurun_out/${TAG}_ln_probe_dbg$d.jsonl; done
python - <<PY
import json
for f in ("bench_default", "bench_reference", "bench_chr1_500kb_alpha_sweep_m20", "bench_all_inputs_a1.1_m20", "bench_synth_10kb_single", "bench_synth_50kb_single"):
    try:
        r = json.loads(open("gpurun_out/${TAG}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, r.get("value"), r.get("ms_per_step"), (r.get("e2e") or {}).get("value"), (r.get("roofline") or {}).get("frac"), r.get("gpu_launches"))
    except Exception as e:
        print(f, "ERR", e)
PY
