mkdir -p gpurun_out
for d in 0 1 2; do echo "dbg=$d"; CHR3D_LN_DEBUG=$d timeout 300 python scripts/large_n_probe.py 24925 2>&1 | grep -v fp32_peak | python -c "
import sys, json
for l in sys.stdin:
    try: r = json.loads(l); print(r['want_energy'], r['tile_ms'], r['frac_hbm'])
    except Exception: print(l.strip())
"; done
