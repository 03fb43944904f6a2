#!/usr/bin/env python3
"""A/B probe of a libchr3d build (CHR3D_LIBRARY): stand-alone tile kernel and whole evaluation of the persistent kernel at N = 24,925 / 4,985."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from chromosome3d_b200 import api
api.init(0)
out = {"lib": os.path.basename(os.environ.get("CHR3D_LIBRARY", "libchr3d.so"))}
for N, seed in ((24925, 20260002), (4985, 20260001)):
    s = api.Single(N=N, alpha=1.1, synth_seed=seed, seed=5)
    t0, e0 = s.eval_bench(reps=20, want_energy=False)
    t1, e1 = s.eval_bench(reps=20, want_energy=True)
    s.profile(); 
    t0b, e0b = s.eval_bench(reps=40, want_energy=False)
    ph = s.profile()
    s.close()
    out[str(N)] = dict(tile_ms=round(t0, 5), eval_ms=round(e0, 5), tile_e_ms=round(t1, 5), eval2_ms=round(e0b, 5),
                       phases_us_per_eval={k: round(1e3 * v / 44.0, 2) for k, v in ph.items()})
print(json.dumps(out), flush=True)
