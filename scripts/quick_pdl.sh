timeout 900 python -m pytest tests/test_gpu_large_n.py -m gpu -x -q 2>&1 | tail -3
for off in 0 1; do
  if [ $off = 1 ]; then export CHR3D_LN_NO_PDL=1; else unset CHR3D_LN_NO_PDL; fi
  echo "NO_PDL=$off"
  timeout 300 python scripts/large_n_probe.py 4985 24925 2>&1 | grep '"want_energy": false' | python -c "
import sys, json
for l in sys.stdin:
    r = json.loads(l); print(r['N'], 'tile_ms', round(r['tile_ms'],4), 'eval_ms', round(r['eval_ms'],4))"
  for w in synth_50kb_single synth_10kb_single; do timeout 300 python bench.py --workload $w --steps 2 --warmup 2 --no-cpu-baseline | python -c "import sys,json; r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w ms', round(r['ms_per_step'],2), r['value'])"; done
done
