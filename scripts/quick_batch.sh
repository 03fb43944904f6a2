mkdir -p gpurun_out
for w in chr1_500kb_a1.1_m20 chr1_500kb_alpha_sweep_m20 all_inputs_a1.1_m20; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-large-n --workload $w > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "exit=$?"
  python - <<PY
import json
r = json.loads(open("gpurun_out/bench_$w.json").read().strip().splitlines()[-1])
print("$w", r["value"], r["ms_per_step"], r.get("sa_models_per_sec"), r["roofline"]["frac"], {k: r["config"].get(k) for k in ("cluster_size", "threads_per_cta", "sa_launches_per_step")})
PY
done
python scripts/img_probe.py 40 1
